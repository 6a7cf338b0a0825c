#!/usr/bin/env python
"""Ingest measurement: text dataset -> dense indices + CSR/CSC.
  this repo : qmf_b200/host/bin/ingest_bench (mapped multi-threaded parser + qmfb_signals_build on the GPU)
  reference : its own DatasetReader::readAll and WALSEngine::init (groupSignals) through oracle/_ref
usage: bench_ingest.py [nnz=20000000] [nusers=480000] [nitems=17800] [--no-ref]"""
import ctypes as C, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
args = [a for a in sys.argv[1:] if not a.startswith("--")]
nnz = int(args[0]) if len(args) > 0 else 20_000_000
nu = int(args[1]) if len(args) > 1 else 480_000
ni = int(args[2]) if len(args) > 2 else 17_800
path = "/tmp/qmf_b200_ingest_%d.txt" % nnz
host = os.path.join(ROOT, "qmf_b200", "host", "bin")
t = time.time()
subprocess.check_call([os.path.join(host, "gen_dataset"), str(nnz), str(nu), str(ni), path, "1"])
out = {"file_bytes": os.path.getsize(path), "gen_s": round(time.time() - t, 2)}
env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "qmf_b200"))
if "--no-ours" not in sys.argv:
    out.update(json.loads(subprocess.check_output([os.path.join(host, "ingest_bench"), path, "0"], env=env).decode().strip().splitlines()[-1]))
if "--no-ref" not in sys.argv:
    import numpy as np
    import oracle
    if oracle.ref_available():
        R = oracle.ref()
        u, i, v = np.empty(nnz, np.int64), np.empty(nnz, np.int64), np.empty(nnz, np.float64)
        t = time.time()
        n = R.ref_read_dataset(path.encode(), u.ctypes.data, i.ctypes.data, v.ctypes.data, nnz)
        out["reference_read_s"] = round(time.time() - t, 3)
        assert n == nnz
        h = R.ref_wals_create(8, 1, 0.05, 40.0, 16, None, 0, 0, 1)
        t = time.time()
        R.ref_wals_init(h, u, i, v, nnz)
        out["reference_init_s"] = round(time.time() - t, 3)  # groupSignals x2 + IdIndex + factor init (k = 8)
        R.ref_wals_destroy(h)
os.remove(path)
if "read_s" in out:
    out["ours_total_s"] = round(out["read_s"] + out["signals_gpu_s"], 3)
if "reference_read_s" in out:
    out["reference_total_s"] = round(out["reference_read_s"] + out["reference_init_s"], 3)
print(json.dumps(out))

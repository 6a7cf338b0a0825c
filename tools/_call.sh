# scratch: the command file tools/gpurun_retry.sh sends to the GPU box (`bash tools/_call.sh`); overwritten per experiment.
# The round's evidence capture is tools/capture_profiles.sh.
set -x
python -c "import __graft_entry__ as g; g.smoke()"

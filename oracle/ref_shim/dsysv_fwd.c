/* TEST INFRASTRUCTURE — the reference links an un-vendored system LAPACK for dsysv_
 * (qmf/Matrix.cpp:25-35,92; CMakeLists.txt:59).  This image has no liblapack; the LP64 OpenBLAS
 * bundled in the scipy wheel (OpenBLAS 0.3.31.dev, scipy.libs/libscipy_openblas-*.so) exports the
 * same routine as scipy_dsysv_.  Forward to it.  */
extern void scipy_dsysv_(char* uplo, int* n, int* nrhs, double* a, int* lda, int* ipiv, double* b, int* ldb,
                         double* work, int* lwork, int* info);
void dsysv_(char* uplo, int* n, int* nrhs, double* a, int* lda, int* ipiv, double* b, int* ldb, double* work,
            int* lwork, int* info) {
  scipy_dsysv_(uplo, n, nrhs, a, lda, ipiv, b, ldb, work, lwork, info);
}

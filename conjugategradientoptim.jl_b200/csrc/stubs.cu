// stubs.cu — entry points declared in include/cgoptim.h that are not implemented yet.
#include "internal.cuh"
#define NOT_YET(name) cgo_set_error(name ": not implemented yet"); return 99
extern "C" int cgo_obj_sparse_ls_create_synthetic(cgo_ctx *, int64_t, int32_t, int64_t, uint64_t, int32_t, cgo_obj **) { NOT_YET("cgo_obj_sparse_ls_create_synthetic"); }
extern "C" int cgo_obj_sparse_ls_create_csr(cgo_ctx *, int64_t, int64_t, const int64_t *, const int32_t *, const double *, const double *, cgo_obj **) { NOT_YET("cgo_obj_sparse_ls_create_csr"); }
extern "C" int cgo_obj_logreg_create_synthetic(cgo_ctx *, int64_t, int64_t, int32_t, uint64_t, double, cgo_obj **) { NOT_YET("cgo_obj_logreg_create_synthetic"); }
extern "C" int cgo_obj_csr_nnz(cgo_obj *, int, int64_t *, int64_t *) { NOT_YET("cgo_obj_csr_nnz"); }
extern "C" int cgo_obj_csr_download(cgo_obj *, int, int64_t *, int32_t *, double *, double *) { NOT_YET("cgo_obj_csr_download"); }
extern "C" int cgo_obj_spmv(cgo_obj *, int, const double *, double *) { NOT_YET("cgo_obj_spmv"); }
extern "C" int cgo_batched_minimize_rosenbrock(cgo_ctx *, int64_t, int32_t, const double *, const cgo_batched_config *, double *, int64_t *, int32_t *, int64_t *, double *, double *) { NOT_YET("cgo_batched_minimize_rosenbrock"); }

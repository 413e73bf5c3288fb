// stubs.cu — entry points declared in include/cgoptim.h that are not implemented yet.
#include "internal.cuh"
#define NOT_YET(name) cgo_set_error(name ": not implemented yet"); return 99
extern "C" int cgo_batched_minimize_rosenbrock(cgo_ctx *, int64_t, int32_t, const double *, const cgo_batched_config *, double *, int64_t *, int32_t *, int64_t *, double *, double *) { NOT_YET("cgo_batched_minimize_rosenbrock"); }

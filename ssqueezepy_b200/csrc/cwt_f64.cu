// float64 / complex128 instantiation of the CWT plan and its kernels.
#include "cwt_impl.cuh"
namespace ssqb {
CwtPlanBase* make_cwt_plan_f64(const ssqb_cwt_desc* d, int* err) { return make_cwt_plan<double>(d, err); }
}

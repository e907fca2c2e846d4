// float32 / complex64 instantiation of the CWT plan and its kernels.
#include "cwt_impl.cuh"
namespace ssqb {
CwtPlanBase* make_cwt_plan_f32(const ssqb_cwt_desc* d, int* err) { return make_cwt_plan<float>(d, err); }
}

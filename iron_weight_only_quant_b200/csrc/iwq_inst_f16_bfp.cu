// explicit instantiation of the F16 / BFP kernels (see iwq_launch.cuh)
#define IWQ_INSTANTIATE
#include "iwq_launch.cuh"
namespace iwq {
IWQ_INSTANTIATE_KIND(IWQ_F16, KIND_BFP)
}

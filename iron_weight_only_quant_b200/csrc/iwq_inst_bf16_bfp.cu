// explicit instantiation of the BF16 / BFP kernels (see iwq_launch.cuh)
#define IWQ_INSTANTIATE
#include "iwq_launch.cuh"
namespace iwq {
IWQ_INSTANTIATE_KIND(IWQ_BF16, KIND_BFP)
}

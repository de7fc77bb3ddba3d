// placeholder until the residual-network kernels land
#include "pipeline.h"
namespace mz {
struct ResNetDevice { int dummy; };
ResNetDevice* resnet_create(const MzNetDesc&, int, int, std::string*) { return new ResNetDevice(); }
void resnet_destroy(ResNetDevice* r) { delete r; }
int resnet_load_weights(ResNetDevice*, const MzTensor*, int, std::string* err) { *err = "resnet kernels not built yet"; return MZ_EUNSUPPORTED; }
int resnet_inference(ResNetDevice*, const InferCall&, cudaStream_t, int64_t*, std::string* err) { *err = "resnet kernels not built yet"; return MZ_EUNSUPPORTED; }
}

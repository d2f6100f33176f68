// model_rocket.cu -- instantiates the batched iLQR kernels for the Rocket model (models.cuh).
#include "engine.cuh"

namespace mas_b200 {
BatchBase* make_batch_rocket() { return new BatchImpl<Rocket>(); }
}  // namespace mas_b200

// model_pendulum.cu -- instantiates the batched iLQR kernels for the Pendulum model (models.cuh).
#include "engine.cuh"

namespace mas_b200 {
BatchBase* make_batch_pendulum() { return new BatchImpl<Pendulum>(); }
}  // namespace mas_b200

// model_build.cu — placeholder until the device model builder lands.
#include "model_build.h"

#include "../../include/itrails_b200.h"

struct ModelPlan {};

ModelBuilder::ModelBuilder() {}
ModelBuilder::~ModelBuilder() { delete plan_; }

int ModelBuilder::build(cudaStream_t, int, const double *, int, int, const double *, const double *,
                        const double **, const double **, const double **, int32_t *, int64_t *launched,
                        std::string &msg) {
    if (launched) *launched = 0;
    msg = "device model builder not available in this build";
    return ITR_ERR_UNSUPPORTED;
}

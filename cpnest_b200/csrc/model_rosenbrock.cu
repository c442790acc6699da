#include "model_inst.cuh"
CPN_DEFINE_MODEL_LAUNCHERS(rosenbrock, cpn::Rosenbrock)

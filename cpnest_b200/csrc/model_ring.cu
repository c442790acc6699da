#include "model_inst.cuh"
CPN_DEFINE_MODEL_LAUNCHERS(ring, cpn::Ring)

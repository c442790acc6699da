#include "model_inst.cuh"
CPN_DEFINE_MODEL_LAUNCHERS(gauss_mix_nd, cpn::GaussMixND)

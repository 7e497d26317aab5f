// sw_params.cuh -- generated field / curve parameter structs (tools/gen_curve_params.py).
#pragma once
#include "fp_mont.cuh"
#include "fp_special.cuh"
namespace eb {
#include "sw_params_gen.inc"
}  // namespace eb

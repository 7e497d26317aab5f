// sw_params.cuh -- generated Montgomery parameter structs (tools/gen_curve_params.py).
#pragma once
#include "fp_mont.cuh"
namespace eb {
#include "sw_params_gen.inc"
}  // namespace eb

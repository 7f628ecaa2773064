// Backward kernels (filled in below the forward path; see DESIGN.md "Backward").
#pragma once
#include "c3d_common.cuh"

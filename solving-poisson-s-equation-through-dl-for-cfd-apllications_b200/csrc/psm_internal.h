// Internal (non-ABI) hooks between the translation units of libpsm_b200.so.
#pragma once
#include "../../include/psm_b200.h"

namespace psm {
int handle_shape(const psm_handle* h);
int handle_variant(const psm_handle* h);
int handle_device(const psm_handle* h);
double handle_delta(const psm_handle* h);                                  // block edge S the handle was created with
int handle_fail(psm_handle* h, int code, const char* fmt, ...);         // sets psm_last_error(h), returns code
}  // namespace psm

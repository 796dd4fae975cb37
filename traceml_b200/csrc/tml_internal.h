// Private glue between the translation units of libtraceml_b200.so (not part of the ABI).
#pragma once

#include "../../include/traceml_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

// sets tml_last_error() for the calling thread and returns `code`
int tml_set_error_(int code, const char* fmt, ...);
// storage for tml_reduce_run's workspace inside the context (owned by tml_summary.cpp)
void** tml_run_ws_slot_(tml_ctx* ctx);
void tml_run_ws_free_(void* ws);

#ifdef __cplusplus
}
#endif

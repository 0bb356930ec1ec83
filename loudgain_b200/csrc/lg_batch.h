// lg_batch.h -- internals shared by the host-layer translation units.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>

#include "lg_common.h"

// The library is built with -fvisibility=hidden; only the C ABI is exported.
#define LG_EXPORT __attribute__((visibility("default")))

namespace lg {

void set_error(const char* what, cudaError_t e);
void set_error(const char* what);

// Runs one gating/range query over the union of `n` block lists (device
// pointers inside) and waits for the result.  0 on success.
int query_lists_sync(const BlockList* lists, size_t n, cudaStream_t stream, QueryResult* out);
// ... and `n` queries at once, query i over list i alone (out[n]).
int query_each_sync(const BlockList* lists, size_t n, cudaStream_t stream, QueryResult* out);

}  // namespace lg

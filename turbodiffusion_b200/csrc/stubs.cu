// Entry points declared in include/tdb200.h whose kernels have not landed yet.  Each returns
// TDB200_ERR_UNSUPPORTED with a message (never a silent fallback).  This file shrinks to nothing as kernels land.
#include "host_common.h"

#define TDB_STUB(name, ...) \
  extern "C" int name(__VA_ARGS__) { return tdb::fail(TDB200_ERR_UNSUPPORTED, #name ": not implemented in this build"); }

TDB_STUB(tdb200_sla_quant_qk, const void*, const void*, int, int64_t, int64_t, int64_t, int64_t, float*, int8_t*,
         float*, int8_t*, float*, void*, void*, void*)
TDB_STUB(tdb200_sla_block_map, const void*, const void*, int, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t,
         int8_t*, int32_t*, void*)
TDB_STUB(tdb200_sla_linear_moments, const void*, const void*, int, int64_t, int64_t, int64_t, int64_t, float*, float*,
         void*)
TDB_STUB(tdb200_sla_attn_fwd, const int8_t*, const float*, const int8_t*, const float*, const void*, const void*, int,
         const int32_t*, int64_t, const void*, const float*, const float*, void*, int64_t, int64_t, int64_t, int64_t,
         int64_t, float, void*)

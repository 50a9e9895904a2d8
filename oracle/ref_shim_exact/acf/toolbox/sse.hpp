// "exact" arithmetic mode for the oracle/_ref build: include the reference's own
// toolbox/sse.hpp unchanged, but route its two approximate intrinsics wrappers
// (RCP -> rcpps, RCPSQRT -> rsqrtps, toolbox/sse.hpp:185-192) to IEEE-correct
// 1/x and 1/sqrt(x). Everything else in the reference TUs compiles as shipped.
#pragma once
#define RCP RCP_approx_shipped
#define RCPSQRT RCPSQRT_approx_shipped
#include REF_SSE_HPP
#undef RCP
#undef RCPSQRT
inline __m128 RCP(const __m128 x) { return _mm_div_ps(_mm_set1_ps(1.0f), x); }
inline __m128 RCPSQRT(const __m128 x) { return _mm_div_ps(_mm_set1_ps(1.0f), _mm_sqrt_ps(x)); }

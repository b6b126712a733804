/*
 * ref_nnue_state_gxx.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Compiles the reference's src/eval/nnue_state.cpp UNMODIFIED, from where it lies, with g++.
 * The AVX-512 VBMI2 branch at nnue_state.cpp:239-240 passes __m512i/__m256i values to
 * _mm512_insertf64x4 (a __m512d intrinsic); clang converts implicitly, g++ refuses.
 * Both intrinsics are the same bit-level lane insert, so after <immintrin.h> has been
 * included we redirect the name to the integer-typed twin and then include the source.
 */
#include <immintrin.h>
#define _mm512_insertf64x4(a, b, imm) _mm512_inserti64x4((a), (b), (imm))
#include "eval/nnue_state.cpp"

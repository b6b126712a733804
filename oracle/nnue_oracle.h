/*
 * nnue_oracle.h -- TEST INFRASTRUCTURE ONLY.
 * Plain-C restatement of Stormphrax's NNUE evaluation path (see nnue_oracle.c).
 * Parity status: PINNED -- checked against the reference's own compiled code
 * (oracle/_ref) in tests/test_oracle.py and against tests/golden/ fixtures generated from it.
 */
#ifndef SP_NNUE_ORACLE_H
#define SP_NNUE_ORACLE_H

#include "../include/sp_types.h"

#ifdef __cplusplus
extern "C" {
#endif

int spo_load_net(const void* image, size_t len);
int spo_eval_once(const SpPackedBoard* boards, size_t n, int32_t* out);
double spo_time_eval_once(const SpPackedBoard* boards, size_t n, int threads, int reps, int32_t* out);
int spo_psq_features(const SpPackedBoard* board, int c, uint32_t* out);
int spo_threat_features(const SpPackedBoard* board, int c, uint32_t* out);
int32_t spo_threat_index(int c, int king_sq, int attacker, int attacker_sq, int attacked, int attacked_sq);
int spo_accumulators(const SpPackedBoard* board, int16_t* psq /*[2][1024]*/, int16_t* thr /*[2][1024]*/);
int spo_ft_activations(const SpPackedBoard* board, uint8_t* out /*[1024]*/, int* bucket);
int32_t spo_forward(const uint8_t* ft /*[1024]*/, int bucket);
int32_t spo_forward_acc(const int16_t* psq, const int16_t* thr, int stm, int bucket);

#ifdef __cplusplus
}
#endif

#endif

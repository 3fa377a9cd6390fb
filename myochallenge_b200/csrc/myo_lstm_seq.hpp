// Persistent tcgen05 recurrent kernels of the PPO update (myo_lstm_seq.cu), called from myo_ppo.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace myo {

// hidden sizes the cluster kernels cover (multiples of 64 up to 256: W_hh slice + A tile fit one SM's shared memory)
bool lstm_seq_supported(int H);
size_t lstm_seq_wpack_bytes(int H);      // packed bf16 image of W_hh (per network, per direction)
constexpr int kLstmRecBytes = 96;

struct LstmSeqFwd {
  int T, B, H;
  const float *whh, *bih, *bhh;           // fp32 parameters: [4H][H], [4H], [4H]
  uint8_t* wpack;                         // scratch: lstm_seq_wpack_bytes(H)
  float* bias;                            // scratch: [4H]
  const float* G;                         // [T][B][4H]: x W_ih^T (input projection)
  const float* keep;                      // [T][B]: 1 - episode_start
  const float* C0;                        // [B][H]: keep_0 c0
  uint8_t* Rec;                           // [T][B][H/8] activation records (kLstmRecBytes each): gates i f g o as 8 x bf16 each, c_t as 8 x fp32
  void* Hs;                               // bf16 [T][B][H]: h_t
  void* HP;                               // bf16 [T][B][H]: HP[0] = keep_0 h0 in; HP[t+1] = keep_{t+1} h_t out
  long long* prof = nullptr;              // optional device [8]: cycle counters of one thread (development)
};
int lstm_seq_forward(const LstmSeqFwd& f, cudaStream_t st);

}  // namespace myo

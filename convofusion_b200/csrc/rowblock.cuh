// Row-block kernel: one persistent, warp-specialised tcgen05 CTA per 128-row block of the denoiser's query side.
// See rowblock.cu for the design; this header is the host-side interface used by denoiser.cu.
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace cfb {

enum RbKind { RB_HLOAD = 0, RB_GEMM = 1 };
enum RbASrc { RB_A_KEEP = 0, RB_A_TMA = 1 };
enum RbEpi {
  RB_EPI_NONE = 0,     // accumulate only (more K follows in the next stage)
  RB_EPI_LN = 1        // residual update complete: + bias, row statistics, LayerNorm (+ TimeBlock modulation + SiLU)
                       // of the updated rows -> next A operand (bf16, shared memory)
};

// One step of a row block's program.  Lives in global memory (tensor maps included, 64-byte aligned).
struct alignas(128) RbStage {
  CUtensorMap map_w;        // weights [N_total, K_total] bf16, box 64 x 128
  CUtensorMap map_a;        // A operand source [rows, K_total] bf16, box 64 x 128 (a_src == RB_A_TMA)
  int kind;                 // RbKind
  int a_src;                // RbASrc
  int K, N;                 // K multiple of 64 (<= 512), N multiple of 128 (<= 512)
  int w_row0, w_col0;       // first W row (= output column block) and first W column (= K offset)
  int a_col0;               // first column of the A source
  int per_stream;           // 1: conditional-stream stage -- skipped for blocks without a conditional stream; W / A
                            // column offsets advance by stream * 512
  int epi;                  // RbEpi
  int ln_silu;              // TimeBlock: LN(h) * (1 + scale) + shift, then SiLU
  int spill;                // write the updated residual rows (fp32) back to global memory
  int store_a;              // write the new A operand (bf16) to global memory
  int commit_a;             // the next GEMM stage loads A by TMA: release every A panel as soon as it has been consumed
  int wait_epi;             // the stage's MMAs wait for the preceding epilogue phase (residual load / LayerNorm)
  int pad0, pad1;
  const float* bias;        // [N] or null
  const float* ln_g;        // [512]
  const float* ln_b;        // [512]
  const float* mod;         // [steps][mod_stride]: scale(512) | shift(512) of this TimeBlock, or null
  long long mod_stride;
};

struct RbLaunch {
  const RbStage* prog;      // device
  int n_stages;
  int block0, n_blocks;     // row blocks [block0, block0 + n_blocks) of 128 rows
  const int* blk_stream;    // device [total blocks]: conditional stream of the block's rows, -1 = none
  const int* step_ptr;      // device step counter (modulation row) or null
  float* h;                 // residual stream [rows_total, 512] fp32
  bf16* a;                  // A operand buffer [rows_total, 512] bf16 (store_a target)
  int rows_total;
  int trace;                // debug: candidate for phase tracing (rowblock_trace_arm)
};

int init_rowblock_kernels();
int rowblock_launch(const RbLaunch& L, cudaStream_t st);
// tensor map of a K-major bf16 operand [rows, cols] with leading dimension ld, box 64 x 128 (TMA load or store)
int rowblock_operand_map(const void* p, int rows, int cols, int ld, CUtensorMap* out);
// host-mapped fault record of the last protocol timeout: {code, block, warp, stage, barrier}; all zero = none
const unsigned* rowblock_fault_record();
void rowblock_trace_arm(int n);
int rowblock_trace_read(unsigned long long out[64]);

}  // namespace cfb

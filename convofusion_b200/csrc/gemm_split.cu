// fp32-accurate GEMMs on the tcgen05 tensor cores: y = A W^T with both fp32 operands split three ways into bf16
// (x = hi + mid + lo exactly, 8 + 8 + 8 mantissa bits) and the six significant partial products
//     hi.hi + hi.mid + mid.hi + hi.lo + mid.mid + lo.hi          (dropped: mid.lo, lo.mid, lo.lo <= 2^-24 |a||w|)
// accumulated in the fp32 tensor-memory accumulator.  bf16 x bf16 products are exact in fp32, so the result carries
// fp32-level error (measured against float64 in tests/test_gpu_kernels.py) at 6x the MMA work of the bf16 mode -- on a
// pipe that is < 25 % busy -- instead of the CUDA-core FFMA GEMM.
//
// No new MMA kernel: the split is laid out so that the EXISTING bf16 kernel (gemm_tc.cu) computes it as one GEMM with
// K' = 6 K.  For every 64-column block of the operand, six 64-column blocks follow each other in the split matrix:
//     A side:  hi  hi  mid hi  mid lo        W side:  hi  mid hi  lo  mid hi
// so that block t of A meets block t of W.  Any 64-aligned K slice of an operand is a contiguous 6x wider slice of its
// split.  Activations are split on the fly into a stream-ordered scratch; weights are split once per handle (cache).
//
// Reduced schemes for the precision study (tools/precision_study.py): scheme 1 keeps the
// activations at two bf16 terms (hi + lo, 16 mantissa bits) against bf16-rounded weights (2 blocks per 64 columns:
// A hi lo / W hi hi), scheme 2 rounds both operands to bf16 (1 block) -- the bf16 mode's GEMM rounding in isolation.
#include "common.cuh"
#include <cstdlib>
#include <mutex>
#include <unordered_map>

namespace cfb {

namespace {

struct Key {
  const void* p; int rows, K; long long ld; int role;   // role = operand side + 2 * scheme
  bool operator==(const Key& o) const { return p == o.p && rows == o.rows && K == o.K && ld == o.ld && role == o.role; }
};
// which term (0 hi, 1 mid, 2 lo) goes into block t of the split operand, per scheme and side
__constant__ int c_comp[3][2][6] = {{{0, 0, 1, 0, 1, 2}, {0, 1, 0, 2, 1, 0}}, {{0, 1, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0}},
                                    {{0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0}}};
constexpr int kBlocks[3] = {6, 2, 1};
struct KeyHash {
  size_t operator()(const Key& k) const {
    size_t h = std::hash<const void*>()(k.p);
    h = h * 1000003u ^ (size_t)k.rows; h = h * 1000003u ^ (size_t)k.K; h = h * 1000003u ^ (size_t)k.ld;
    return h * 1000003u ^ (size_t)k.role;
  }
};

// side 0 = A operand, 1 = W operand; `nblk` blocks of 64 columns per 64 input columns (6 / 2 / 1 by scheme)
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ X, long long ld, bf16* __restrict__ out,
                                                     int rows, int K, int scheme, int side, int nblk) {
  pdl_sync();
  const int c8n = K >> 3;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * c8n) return;
  const int r = (int)(idx / c8n), c = (int)(idx % c8n) * 8;
  const float4 x0 = *reinterpret_cast<const float4*>(X + (size_t)r * ld + c);
  const float4 x1 = *reinterpret_cast<const float4*>(X + (size_t)r * ld + c + 4);
  const float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
  __align__(16) bf16 term[3][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    term[0][i] = __float2bfloat16_rn(x[i]);
    const float r1 = x[i] - __bfloat162float(term[0][i]);            // exact
    term[1][i] = __float2bfloat16_rn(r1);
    term[2][i] = __float2bfloat16_rn(r1 - __bfloat162float(term[1][i]));  // exact difference, rounded once
  }
  bf16* o = out + (size_t)r * nblk * K + (size_t)(c >> 6) * (64 * nblk) + (c & 63);
  for (int t = 0; t < nblk; ++t)
    *reinterpret_cast<uint4*>(o + 64 * t) = *reinterpret_cast<const uint4*>(term[c_comp[scheme][side][t]]);
}

int split3(const float* X, long long ld, bf16* out, int rows, int K, int side, int scheme, cudaStream_t st) {
  const long long n = (long long)rows * (K >> 3);
  if (n <= 0) return CFB_OK;
  const unsigned grid = (unsigned)((n + 255) / 256);
  launch_k(split3_kernel, dim3(grid), dim3(256), 0, st, X, ld, out, rows, K, scheme, side, kBlocks[scheme]);
  CFB_LAUNCH_CHECK();
  return CFB_OK;
}

}  // namespace

struct SplitCache {
  std::mutex mu;
  std::unordered_map<Key, void*, KeyHash> map;
  size_t bytes = 0;
};

SplitCache* split_cache_create() { return new SplitCache(); }
void split_cache_destroy(SplitCache* c) {
  if (!c) return;
  for (auto& kv : c->map) cudaFree(kv.second);
  delete c;
}
size_t split_cache_bytes(const SplitCache* c) { return c ? c->bytes : 0; }

// Split of a STATIC operand (a weight): computed on first use, kept for the lifetime of the cache.
int split_static(SplitCache* c, const float* X, int rows, int K, long long ld, int side, int scheme, cudaStream_t st,
                 const bf16** out) {
  CFB_CHECK(c != nullptr, "gemm_split: static operand without a cache");
  std::lock_guard<std::mutex> g(c->mu);
  const Key key{X, rows, K, ld, side + 2 * scheme};
  auto it = c->map.find(key);
  if (it != c->map.end()) { *out = reinterpret_cast<const bf16*>(it->second); return CFB_OK; }
  CFB_CHECK(!t_capturing, "gemm_split: a weight would be split inside a stream capture (prefill the cache first)");
  void* p = nullptr;
  const size_t bytes = (size_t)rows * kBlocks[scheme] * K * 2;
  CFB_CUDA(cudaMalloc(&p, bytes));
  c->bytes += bytes;
  c->map.emplace(key, p);
  CFB_TRY(split3(X, ld, reinterpret_cast<bf16*>(p), rows, K, side, scheme, st));
  *out = reinterpret_cast<const bf16*>(p);
  return CFB_OK;
}

bool gemm_split_supported(int M, int N, int K, long long lda, long long ldw) {
  return K % 64 == 0 && lda % 4 == 0 && ldw % 4 == 0 && gemm_tc_supported(M, N, 6 * K, 6 * K, 6 * K);
}

int gemm_split(const float* A, long long lda, const float* W, long long ldw, int M, int N, int K, const Epilogue& ep,
               cudaStream_t st) {
  const SplitCtx* sc = ep.split;
  CFB_CHECK(sc != nullptr && gemm_split_supported(M, N, K, lda, ldw), "gemm_split: unsupported call %dx%dx%d", M, N, K);
  CFB_CHECK(((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0), "gemm_split: operands must be 16-byte aligned");
  const bf16 *As = nullptr, *Ws = nullptr;
  // scheme 3 (precision study): hi + lo activations only where the A operand is a LayerNorm output, bf16 elsewhere
  static const int sites = getenv("CFB_SPLIT_SITES") ? atoi(getenv("CFB_SPLIT_SITES")) : 31;   // which LayerNorm-fed GEMMs
  const int scheme = sc->scheme == 3 ? ((ep.a_from_ln & sites) ? 1 : 2) : sc->scheme;
  CFB_CHECK(scheme >= 0 && scheme < 3, "gemm_split: unknown scheme %d", scheme);
  const int nb = kBlocks[scheme];
  if (ep.a_static) {
    CFB_TRY(split_static(sc->cache, A, M, K, lda, 0, scheme, st, &As));
  } else {
    CFB_CHECK((size_t)M * nb * K * 2 <= sc->a_ws_bytes, "gemm_split: A scratch too small (%d x %d)", M, K);
    CFB_TRY(split3(A, lda, reinterpret_cast<bf16*>(sc->a_ws), M, K, 0, scheme, st));
    As = reinterpret_cast<const bf16*>(sc->a_ws);
  }
  if (ep.w_static) {
    CFB_TRY(split_static(sc->cache, W, N, K, ldw, 1, scheme, st, &Ws));
  } else {
    CFB_CHECK((size_t)N * nb * K * 2 <= sc->w_ws_bytes, "gemm_split: W scratch too small (%d x %d)", N, K);
    CFB_TRY(split3(W, ldw, reinterpret_cast<bf16*>(sc->w_ws), N, K, 1, scheme, st));
    Ws = reinterpret_cast<const bf16*>(sc->w_ws);
  }
  Epilogue e2 = ep;
  e2.split = nullptr;
  return gemm_tc(As, nb * K, Ws, nb * K, M, N, nb * K, e2, st);
}

}  // namespace cfb

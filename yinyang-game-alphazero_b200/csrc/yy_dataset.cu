// yy_dataset.cu -- replay records -> training tensors with the 8-fold symmetry augmentation, on the device.
//
// Restates create_dataset_from_games / DataProcessor.preprocess_sample / augment_sample
// (src/yin_yang/ai/data_utils.py:16-215) over board_to_input (src/yin_yang/ai/neural_network.py:156-196) and
// Node.get_children_distribution at temperature 1 (src/yin_yang/ai/mcts.py:183-215) for a whole replay buffer:
// for every record the input planes [5][n][m] (empty, black, white, row fill, column fill) and the policy [A] are
// written in the 8 forms of augment_sample, in its order (identity, rot90 x1/x2/x3 counter-clockwise, flip left-right,
// flip up-down, transpose, anti-transpose), the value replicated.  All arithmetic is the reference's: fractions are
// float64 quotients rounded to float32 (Python float -> torch.float32), the forms are pure permutations => bit-exact.
//
// Roofline: HBM, write bound.  Algorithmic bytes per record = 16W + 2A + 4 read, 8*(6A + 1)*4 written
// (12,452 B at 8x8).  One warp per record; every store instruction writes 32 consecutive floats.
#include "yy_common.cuh"

namespace yy {

constexpr int kDatasetBlock = 256;   // 8 records per CTA

template <int NW>
__global__ void __launch_bounds__(kDatasetBlock)
augment_kernel(Geo<NW> g, int W, const uint64_t* __restrict__ black, const uint64_t* __restrict__ white,
               const uint16_t* __restrict__ counts, const float* __restrict__ policy_in, const float* __restrict__ values,
               long long count, float* __restrict__ out_planes, float* __restrict__ out_policy, float* __restrict__ out_values) {
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= count) return;
  const int n = g.rows, m = g.cols, A = g.cells;
  const BB<NW> b = load_bb<NW>(black, r, W) & g.full, w = load_bb<NW>(white, r, W) & g.full;
  const BB<NW> occ = b | w;
  // row / column fill fractions (neural_network.py:183-194): lane x holds row x's, lane y column y's  (n, m <= 32)
  int rc = 0, cc = 0;
  for (int y = 0; y < m; ++y) rc += (lane < n && test(occ, lane * m + y)) ? 1 : 0;
  for (int x = 0; x < n; ++x) cc += (lane < m && test(occ, x * m + lane)) ? 1 : 0;
  const float rowf = (float)((double)rc / (double)m), colf = (float)((double)cc / (double)n);
  // policy = visit counts / their sum in float64 (uniform when there is no visit, mcts.py:209-213), then float32
  unsigned total = 0;
  if (counts) {
    for (int a = lane; a < A; a += 32) total += counts[r * A + a];
    for (int off = 16; off; off >>= 1) total += __shfl_xor_sync(0xffffffffu, total, off);
  }
  const float uniform = (float)(1.0 / (double)A);
  const float v = values ? values[r] : 0.0f;
  if (lane < 8 && out_values) out_values[r * 8 + lane] = v;
  for (int f = 0; f < 8; ++f) {
    float* planes = out_planes + (r * 8 + f) * 5ll * A;
    float* pol = out_policy + (r * 8 + f) * (long long)A;
    for (int a0 = 0; a0 < A; a0 += 32) {
      const int a = a0 + lane;
      const bool live = a < A;
      const int i = live ? a / m : 0, j = live ? a - i * m : 0;
      int sx, sy;   // source cell of output cell (i, j)   (square boards: n == m)
      switch (f) {
        case 0: sx = i; sy = j; break;
        case 1: sx = j; sy = m - 1 - i; break;             // np.rot90(S, 1)[i][j] = S[j][m-1-i]
        case 2: sx = n - 1 - i; sy = m - 1 - j; break;     // rot90 x2
        case 3: sx = n - 1 - j; sy = i; break;             // rot90 x3
        case 4: sx = i; sy = m - 1 - j; break;             // flip left-right
        case 5: sx = n - 1 - i; sy = j; break;             // flip up-down
        case 6: sx = j; sy = i; break;                     // transpose
        default: sx = n - 1 - j; sy = m - 1 - i; break;    // flip(transpose, both axes)
      }
      const int s = sx * m + sy;
      const float rf = __shfl_sync(0xffffffffu, rowf, sx), cf = __shfl_sync(0xffffffffu, colf, sy);
      if (live) {
        const bool isb = test(b, s), isw = test(w, s);
        planes[a] = (isb || isw) ? 0.0f : 1.0f;
        planes[A + a] = isb ? 1.0f : 0.0f;
        planes[2 * A + a] = isw ? 1.0f : 0.0f;
        planes[3 * A + a] = rf;
        planes[4 * A + a] = cf;
        float p;
        if (counts) p = total ? (float)((double)counts[r * A + s] / (double)total) : uniform;
        else p = policy_in[r * A + s];
        pol[a] = p;
      }
    }
  }
}

}  // namespace yy

using namespace yy;

extern "C" int yy_augment_samples(int rows, int cols, const uint64_t* black, const uint64_t* white, const uint16_t* counts,
                                  const float* policy, const float* values, int64_t count, float* out_planes,
                                  float* out_policy, float* out_values, void* stream) {
  if (!board_supported(rows, cols)) return set_error(YY_ERR_INVALID, "unsupported board %dx%d", rows, cols);
  if (rows != cols) return set_error(YY_ERR_INVALID, "augmentation rotates by 90 degrees: square boards only (data_utils.py:59-76)");
  if (count < 0) return set_error(YY_ERR_INVALID, "negative count");
  if (yy_device_count() == 0) return set_error(YY_ERR_NO_DEVICE, "no CUDA device: the engine has no CPU fallback");
  if (count == 0) return YY_OK;     // empty replay buffer: nothing to write (pointers may be null)
  if (!black || !white || (!counts && !policy) || !out_planes || !out_policy) return set_error(YY_ERR_INVALID, "null argument");
  const int cells = rows * cols, W = words_for_cells(cells);
  const unsigned grid = (unsigned)((count * 32 + kDatasetBlock - 1) / kDatasetBlock);
  YY_DISPATCH_NW(cells, augment_kernel<NW><<<grid, kDatasetBlock, 0, (cudaStream_t)stream>>>(
      make_geo<NW>(rows, cols, 0), W, black, white, counts, policy, values, count, out_planes, out_policy, out_values));
  YY_LAUNCH_CHECK();
  return YY_OK;
}

// Kernel-target alignment (row f3 of SURVEY.md §8): the unsupervised accuracy score the reference computes per
// (k, mixture) to choose the mixture parameter (scRNA/utils.py:66-152, cmd_target.py:254-260).
//
// The reference centres both n x n kernels with three dense n^3 products each (center_kernel, utils.py:79-83) and
// takes the Frobenius inner products as traces of further n^3 products (kta_align_general, :86-95).  Both are
// closed forms of the row means:  Kc_ij = K_ij - m_i - m_j + m,  Kn_ij = Kc_ij * (b_i b_j),  b_i = Kc_ii^-1/2,
// so after the linear kernel Kx = X^T X (scrna_pairwise_f64, op 1) the score is two O(n^2) passes:
//   kta_rowstats : row sums and diagonal of Kx (one warp per row, fixed order)
//   kta_align    : sum_ij Kxn_ij Kyn_ij, sum Kxn^2, sum Kyn^2 with Ky_ij = [l_i == l_j] formed on the fly
#include "common.cuh"

namespace scrna {

__global__ void __launch_bounds__(256)
kta_rowstats_kernel(const double* __restrict__ K, int64_t ld, int n, double* __restrict__ rowsum,
                    double* __restrict__ diag) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  double s = 0.0;
  for (int j = lane; j < n; j += 32) s += K[(int64_t)row * ld + j];
  s = warp_sum(s);
  if (lane == 0) {
    rowsum[row] = s;
    diag[row] = K[(int64_t)row * ld + row];
  }
}

// px / py: per-row {mean, scale} of the x and y kernels (scale 1 when not normalised, mean 0 when not centred);
// tot_x / tot_y: the total means.  partials: gridDim.x x 3.
__global__ void __launch_bounds__(256)
kta_align_kernel(const double* __restrict__ K, int64_t ld, int n, const int* __restrict__ labels,
                 const double* __restrict__ mean_x, const double* __restrict__ scale_x, double tot_x,
                 const double* __restrict__ mean_y, const double* __restrict__ scale_y, double tot_y,
                 double* __restrict__ partials) {
  __shared__ double scratch[32];
  double sxy = 0.0, sxx = 0.0, syy = 0.0;
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    const double mxi = mean_x[i], bxi = scale_x[i], myi = mean_y[i], byi = scale_y[i];
    const int li = labels[i];
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      const double kx = (K[(int64_t)i * ld + j] - mxi - mean_x[j] + tot_x) * (bxi * scale_x[j]);
      const double ky = ((labels[j] == li ? 1.0 : 0.0) - myi - mean_y[j] + tot_y) * (byi * scale_y[j]);
      sxy = fma(kx, ky, sxy);
      sxx = fma(kx, kx, sxx);
      syy = fma(ky, ky, syy);
    }
  }
  const double a = block_sum(sxy, scratch);
  const double b = block_sum(sxx, scratch);
  const double c = block_sum(syy, scratch);
  if (threadIdx.x == 0) {
    partials[blockIdx.x * 3 + 0] = a;
    partials[blockIdx.x * 3 + 1] = b;
    partials[blockIdx.x * 3 + 2] = c;
  }
}

// Gaussian kernel from the linear Gram, in place: K_ij <- exp(-(K_ii - 2 K_ij + K_jj) / param)  (get_kernel(type='rbf'),
// scRNA/utils.py:108-124).  The diagonal is read from a copy taken before the pass.
__global__ void __launch_bounds__(256)
kta_rbf_kernel(double* __restrict__ K, int64_t ld, int n, const double* __restrict__ diag, double param) {
  const int i = blockIdx.y;
  const double di = diag[i];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const double v = di - 2.0 * K[(int64_t)i * ld + j] + diag[j];
    K[(int64_t)i * ld + j] = exp(-v / param);
  }
}

__global__ void kta_final_kernel(const double* __restrict__ partials, int nblk, double* __restrict__ out) {
  if (threadIdx.x < 3) {
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += partials[b * 3 + threadIdx.x];  // fixed order
    out[threadIdx.x] = s;
  }
}

}  // namespace scrna

using namespace scrna;

extern "C" int scrna_kta_blocks(void) { return 4 * num_sms(); }

extern "C" int scrna_kta_rowstats_f64(const double* K, int64_t ld, int n, double* rowsum, double* diag, void* stream) {
  if (!K || !rowsum || !diag || n <= 0 || ld < n) return SCRNA_ERR_BADARG;
  kta_rowstats_kernel<<<(unsigned)ceil_div(n, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(K, ld, n, rowsum, diag);
  return last_launch_status();
}

extern "C" int scrna_kta_rbf_f64(double* K, int64_t ld, int n, const double* diag, double param, void* stream) {
  if (!K || !diag || n <= 0 || ld < n || n > 65535 || !(param != 0.0)) return SCRNA_ERR_BADARG;
  dim3 grid((unsigned)(ceil_div(n, 256) > 32 ? 32 : ceil_div(n, 256)), (unsigned)n);
  kta_rbf_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(K, ld, n, diag, param);
  return last_launch_status();
}

extern "C" int scrna_kta_align_f64(const double* K, int64_t ld, int n, const int32_t* labels, const double* mean_x,
                                   const double* scale_x, double tot_x, const double* mean_y, const double* scale_y,
                                   double tot_y, double* partials, double* out, void* stream) {
  if (!K || !labels || !mean_x || !scale_x || !mean_y || !scale_y || !partials || !out || n <= 0 || ld < n)
    return SCRNA_ERR_BADARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int nblk = scrna_kta_blocks();
  if (nblk > n) nblk = n;
  kta_align_kernel<<<nblk, 256, 0, st>>>(K, ld, n, labels, mean_x, scale_x, tot_x, mean_y, scale_y, tot_y, partials);
  kta_final_kernel<<<1, 32, 0, st>>>(partials, nblk, out);
  return last_launch_status();
}

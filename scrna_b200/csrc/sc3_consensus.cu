// SC3 consensus for sm_100a (scRNA/sc3_clustering_impl.py:222-260, scRNA/sc3_clustering.py:88-109).
//
//   K18 coassoc          : integer co-association counts M[i][j] += (label_i == label_j) per labeling, and
//                          consensus = M / cnt (exactly the fp64 values the reference accumulates)
//   K19 row distances    : scrna_pairwise_f64(op=euclid) on the (symmetric) consensus matrix == pdist(rows)
//   K20 nnchain_complete : scipy.cluster.hierarchy.complete == nn_chain(method='complete'), one persistent
//                          CTA; SciPy's tie-breaking (strict '<' scan in index order, previous chain element
//                          preferred) is reproduced because ties are the norm here
//   K21 cut_tree         : stable sort by distance, SciPy's union-find labelling, the BFS/insort order of
//                          _order_cluster_tree, cut at n_clusters, labels numbered by smallest member
#include "common.cuh"

namespace scrna {

__global__ void coassoc_kernel(const int* __restrict__ labels, int n, int* __restrict__ M, int64_t ld) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= n || i >= n) return;
  M[(int64_t)i * ld + j] += (labels[i] == labels[j]);
}

__global__ void consensus_scale_kernel(const int* __restrict__ M, int64_t ldm, int n, double cnt,
                                       double* __restrict__ C, int64_t ldc) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= n || i >= n) return;
  C[(int64_t)i * ldc + j] = (double)M[(int64_t)i * ldm + j] / cnt;
}

// ---- K20 -------------------------------------------------------------------------------------------
struct MinIdx {
  double v;
  int i;
};
__device__ __forceinline__ MinIdx min_idx(MinIdx a, MinIdx b) {
  // smaller value wins; equal values -> smaller index (first occurrence in a scan by index)
  if (b.i >= 0 && (a.i < 0 || b.v < a.v || (b.v == a.v && b.i < a.i))) return b;
  return a;
}

__global__ void __launch_bounds__(1024)
nnchain_complete_kernel(double* __restrict__ D, int64_t ld, int n, int* __restrict__ size,
                        int* __restrict__ chain, double* __restrict__ merges) {
  __shared__ double sv[32];
  __shared__ int si[32];
  __shared__ int s_x, s_y, s_len;
  __shared__ double s_cur;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  for (int i = tid; i < n; i += blockDim.x) size[i] = 1;
  if (tid == 0) s_len = 0;
  __syncthreads();

  for (int k = 0; k < n - 1; ++k) {
    if (s_len == 0) {
      // first active cluster
      int first = n;
      for (int i = tid; i < n; i += blockDim.x)
        if (size[i] > 0) { first = i; break; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
      if (lane == 0) si[warp] = first;
      __syncthreads();
      if (tid == 0) {
        int f = n;
        for (int w = 0; w < nw; ++w) f = min(f, si[w]);
        chain[0] = f;
        s_len = 1;
      }
      __syncthreads();
    }
    while (true) {
      if (tid == 0) {
        const int len = s_len;
        s_x = chain[len - 1];
        if (len > 1) { s_y = chain[len - 2]; s_cur = D[(int64_t)s_x * ld + s_y]; }
        else { s_y = -1; s_cur = INFINITY; }
      }
      __syncthreads();
      const int x = s_x;
      const double* row = D + (int64_t)x * ld;
      MinIdx m{INFINITY, -1};
      for (int i = tid; i < n; i += blockDim.x) {
        if (size[i] == 0 || i == x) continue;
        const double v = row[i];
        if (m.i < 0 || v < m.v) { m.v = v; m.i = i; }   // ascending i per thread: keeps the first minimum
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        MinIdx other{__shfl_xor_sync(0xffffffffu, m.v, o), __shfl_xor_sync(0xffffffffu, m.i, o)};
        m = min_idx(m, other);
      }
      if (lane == 0) { sv[warp] = m.v; si[warp] = m.i; }
      __syncthreads();
      if (tid == 0) {
        MinIdx b{INFINITY, -1};
        for (int w = 0; w < nw; ++w) b = min_idx(b, MinIdx{sv[w], si[w]});
        int y = s_y;
        double cur = s_cur;
        if (b.i >= 0 && b.v < cur) { cur = b.v; y = b.i; }  // strict '<': a tie keeps the previous element
        const int len = s_len;
        if (len > 1 && y == chain[len - 2]) {
          s_y = y; s_cur = cur; s_len = -len;   // negative length signals "reciprocal pair found"
        } else {
          chain[len] = y;
          s_len = len + 1;
        }
      }
      __syncthreads();
      if (s_len < 0) break;
    }
    int x = s_x, y = s_y;
    const double cur = s_cur;
    __syncthreads();
    if (tid == 0) s_len = -s_len - 2;
    if (x > y) { const int t = x; x = y; y = t; }
    const int nx = size[x], ny = size[y];
    __syncthreads();
    if (tid == 0) {
      merges[3 * (int64_t)k + 0] = (double)x;
      merges[3 * (int64_t)k + 1] = (double)y;
      merges[3 * (int64_t)k + 2] = cur;
      size[x] = 0;
      size[y] = nx + ny;
    }
    __syncthreads();
    // Lance-Williams for complete linkage: d(i, x u y) = max(d(i,x), d(i,y)); the merged cluster keeps slot y
    const double* rx = D + (int64_t)x * ld;
    double* ry = D + (int64_t)y * ld;
    for (int i = tid; i < n; i += blockDim.x) {
      if (size[i] == 0 || i == y) continue;
      const double v = fmax(rx[i], ry[i]);
      ry[i] = v;
      D[(int64_t)i * ld + y] = v;
    }
    __syncthreads();
  }
}

// position of element i in a stable ascending sort by (key, tie): pos = #{j : key_j < key_i or
// (key_j == key_i and tie_j < tie_i)}
__global__ void stable_rank_kernel(const double* __restrict__ key, int64_t kstride, const int* __restrict__ tie,
                                   int m, int* __restrict__ pos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double ki = key[(int64_t)i * kstride];
  const int ti = tie ? tie[i] : i;
  int p = 0;
  for (int j = 0; j < m; ++j) {
    const double kj = key[(int64_t)j * kstride];
    const int tj = tie ? tie[j] : j;
    p += (kj < ki) || (kj == ki && tj < ti);
  }
  pos[i] = p;
}

// Build SciPy's linkage matrix Z (sorted by distance, cluster ids by union-find) and the BFS order of
// _order_cluster_tree.  Single thread: O(n alpha(n)) pointer chasing.
//   merges: (n-1) x 3 in NN-chain order, pos: their stable-sorted positions
//   Z: (n-1) x 4 out; negbfs: out, -(BFS index) of the node created by sorted merge i
//   work: ints, 4*(2n-1): parent | left | right | queue
__global__ void linkage_tree_kernel(const double* __restrict__ merges, const int* __restrict__ pos, int n,
                                    double* __restrict__ Z, int* __restrict__ negbfs, int* __restrict__ work) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int total = 2 * n - 1;
  int* parent = work;
  int* left = work + total;
  int* right = work + 2 * total;
  int* queue = work + 3 * total;
  int* sz = negbfs;  // reuse as cluster sizes during the build (n-1 entries are enough: internal nodes)
  for (int i = 0; i < total; ++i) parent[i] = -1;
  // scatter to sorted order
  for (int m = 0; m < n - 1; ++m) {
    const int p = pos[m];
    Z[4 * (int64_t)p + 0] = merges[3 * (int64_t)m + 0];
    Z[4 * (int64_t)p + 1] = merges[3 * (int64_t)m + 1];
    Z[4 * (int64_t)p + 2] = merges[3 * (int64_t)m + 2];
  }
  for (int i = 0; i < n - 1; ++i) {
    int x = (int)Z[4 * (int64_t)i + 0], y = (int)Z[4 * (int64_t)i + 1];
    while (parent[x] >= 0) x = parent[x];
    while (parent[y] >= 0) y = parent[y];
    const int id = n + i;
    const int lo = x < y ? x : y, hi = x < y ? y : x;
    left[id] = lo; right[id] = hi;
    parent[x] = id; parent[y] = id;
    const int sx = x < n ? 1 : sz[x - n], sy = y < n ? 1 : sz[y - n];
    sz[i] = sx + sy;
    Z[4 * (int64_t)i + 0] = (double)lo;
    Z[4 * (int64_t)i + 1] = (double)hi;
    Z[4 * (int64_t)i + 3] = (double)(sx + sy);
  }
  // BFS from the root: pop left of the deque, push right child then left child
  int head = 0, tail = 0, b = 0;
  queue[tail++] = total - 1;
  while (head < tail) {
    const int node = queue[head++];
    if (node >= n) {
      negbfs[node - n] = -(b++);   // sizes are no longer needed for already-built nodes
      queue[tail++] = right[node];
      queue[tail++] = left[node];
    }
  }
}

// Cut: nodes whose position in the (dist asc, BFS desc) order is < n - n_clusters are merged; labels are
// numbered by ascending smallest member (SciPy cut_tree).  Single thread.
__global__ void cut_labels_kernel(const double* __restrict__ Z, const int* __restrict__ order_pos, int n,
                                  int n_clusters, int* __restrict__ labels, int* __restrict__ work) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int total = 2 * n - 1;
  int* parent = work;          // union-find over cluster ids, merged nodes only
  int* lab = work + total;     // label per root
  for (int i = 0; i < total; ++i) { parent[i] = -1; lab[i] = -1; }
  const int take = n - n_clusters;
  for (int i = 0; i < n - 1; ++i) {
    if (order_pos[i] < take) {
      const int id = n + i;
      parent[(int)Z[4 * (int64_t)i + 0]] = id;
      parent[(int)Z[4 * (int64_t)i + 1]] = id;
    }
  }
  int next = 0;
  for (int i = 0; i < n; ++i) {
    int r = i;
    while (parent[r] >= 0) r = parent[r];
    if (lab[r] < 0) lab[r] = next++;
    labels[i] = lab[r];
  }
}

}  // namespace scrna

using namespace scrna;

extern "C" int scrna_coassoc_accum(const int32_t* labels, int n, int32_t* M, int64_t ld, void* stream) {
  if (!labels || !M || n <= 0 || ld < n || n > 65535) return SCRNA_ERR_BADARG;
  dim3 grid((unsigned)ceil_div(n, 128), (unsigned)n);
  coassoc_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(labels, n, M, ld);
  return last_launch_status();
}

extern "C" int scrna_consensus_scale(const int32_t* M, int64_t ldm, int n, double cnt, double* C, int64_t ldc,
                                     void* stream) {
  if (!M || !C || n <= 0 || ldm < n || ldc < n || n > 65535 || cnt <= 0) return SCRNA_ERR_BADARG;
  dim3 grid((unsigned)ceil_div(n, 128), (unsigned)n);
  consensus_scale_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(M, ldm, n, cnt, C, ldc);
  return last_launch_status();
}

extern "C" int scrna_hclust_complete(double* D, int64_t ld, int n, int n_clusters, int32_t* labels, double* Z,
                                     double* merges, int32_t* iwork, void* stream) {
  // D: n x n distances (destroyed); Z: (n-1) x 4 out (SciPy linkage matrix); merges: (n-1) x 3 scratch;
  // iwork: ints, 8*n + 4*(2n-1)
  if (!D || !labels || !Z || !merges || !iwork || n < 2 || ld < n || n_clusters < 1 || n_clusters > n)
    return SCRNA_ERR_BADARG;
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* size = iwork;
  int32_t* chain = iwork + n;
  int32_t* pos = iwork + 2 * (int64_t)n;
  int32_t* negbfs = iwork + 3 * (int64_t)n;
  int32_t* pos2 = iwork + 4 * (int64_t)n;
  int32_t* tree = iwork + 8 * (int64_t)n;
  nnchain_complete_kernel<<<1, 1024, 0, st>>>(D, ld, n, size, chain, merges);
  stable_rank_kernel<<<(unsigned)ceil_div(n - 1, 128), 128, 0, st>>>(merges + 2, 3, nullptr, n - 1, pos);
  linkage_tree_kernel<<<1, 32, 0, st>>>(merges, pos, n, Z, negbfs, tree);
  stable_rank_kernel<<<(unsigned)ceil_div(n - 1, 128), 128, 0, st>>>(Z + 2, 4, negbfs, n - 1, pos2);
  cut_labels_kernel<<<1, 32, 0, st>>>(Z, pos2, n, n_clusters, labels, tree);
  return last_launch_status();
}

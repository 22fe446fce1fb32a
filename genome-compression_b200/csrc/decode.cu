// decode.cu — DAG expansion: sequential decode, batched random access, width.
//
// Replaces shared_tree::iterator (reference src/shared_tree.cpp:553-614), access_leaf
// (:231-236), operator[] (:268-291) and children()/width() (:252-259,
// include/shared_tree.h:165).  The reference walks one explicit DFS stack; here the
// DAG is expanded top-down one whole level per launch: the pointer array of node
// layer k becomes the pointer array of layer k-1 (a mirrored parent swaps its
// children, flags compose downwards), restricted to the positions that cover the
// requested leaf range, so decoding `count` leaves costs O(count + depth).
#include <algorithm>

#include "staging.cuh"
#include "tree.h"

namespace stb {

constexpr int DEC_THREADS = 256;
constexpr int MAX_LAYERS = 40;

struct LayerPtrs {
  const uint2* nodes[MAX_LAYERS];
};

// access_leaf (src/shared_tree.cpp:231-236): stored leaf -> mirrored? -> transposed?
__device__ __forceinline__ unsigned long long apply_leaf(unsigned long long v, uint32_t p, int S) {
  if (p & MIRROR) v = leaf_mirrored(v, S);
  if (p & TRANSPOSE) v = leaf_transposed(v);
  return v;
}

// One expansion step.  cur[] holds the pointers of node layer `k` for positions
// [lo_cur, ...]; writes the child pointers for positions [lo_next, lo_next + n_next).
__global__ void __launch_bounds__(DEC_THREADS)
expand_kernel(const uint32_t* __restrict__ cur, unsigned long long lo_cur, const uint2* __restrict__ nodes,
              uint32_t* __restrict__ next, unsigned long long lo_next, unsigned long long n_next) {
  const unsigned long long j = (unsigned long long)blockIdx.x * DEC_THREADS + threadIdx.x;
  if (j >= n_next) return;
  const unsigned long long c = lo_next + j;  // child position
  const uint32_t p = cur[(c >> 1) - lo_cur];
  uint32_t out = PTR_NULL;
  if (!ptr_is_null(p)) {
    const uint2 nd = __ldg(nodes + (p & IDX_MASK));
    const uint32_t m = (p >> 29) & 1u, t = (p >> 30) & 1u;
    const uint32_t child = (((uint32_t)c & 1u) ^ m) ? nd.y : nd.x;  // mirrored parent: right first (:606-612)
    if (!ptr_is_null(child)) out = compose(child, m, t);            // nulls are skipped before composing
  }
  next[j] = out;
}

__global__ void __launch_bounds__(DEC_THREADS)
leaves_out_kernel(const uint32_t* __restrict__ cur, unsigned long long n, const unsigned long long* __restrict__ leaves,
                  int S, unsigned long long* __restrict__ out) {
  const unsigned long long j = (unsigned long long)blockIdx.x * DEC_THREADS + threadIdx.x;
  if (j >= n) return;
  const uint32_t p = cur[j];
  out[j] = ptr_is_null(p) ? 0ull : apply_leaf(__ldg(leaves + (p & IDX_MASK)), p, S);
}

constexpr int ASCII_TILE = 1024;
__global__ void __launch_bounds__(DEC_THREADS)
ascii_out_kernel(const uint32_t* __restrict__ cur, unsigned long long n, const unsigned long long* __restrict__ leaves,
                 int S, char* __restrict__ out) {
  __shared__ __align__(16) uint8_t stage[ASCII_TILE * 16 + 16];
  const unsigned long long tile_first = (unsigned long long)blockIdx.x * ASCII_TILE;
  const uint32_t here = (uint32_t)min((unsigned long long)ASCII_TILE, n - tile_first);
  const unsigned long long dst = tile_first * (unsigned long long)S;
  const uint32_t shift = (uint32_t)((reinterpret_cast<uintptr_t>(out) + dst) & 3ull);
  // from_nac (src/dna.cpp:51-74), indexed by the 4-bit code
  const char* letters = "SACRGBNKTWVDYHM-";
  for (uint32_t j = threadIdx.x; j < here; j += DEC_THREADS) {
    const uint32_t p = cur[tile_first + j];
    const unsigned long long v = ptr_is_null(p) ? 0ull : apply_leaf(__ldg(leaves + (p & IDX_MASK)), p, S);
    uint8_t* o = stage + shift + j * S;
    for (int i = 0; i < S; ++i) o[i] = (uint8_t)letters[(v >> (4 * i)) & 0xf];
  }
  __syncthreads();
  copy_out_staged<DEC_THREADS>(out + dst - shift, stage, shift, here * S);
}

// operator[] batched: one thread per query, depth dependent gathers.
__global__ void __launch_bounds__(DEC_THREADS)
random_access_kernel(const unsigned long long* __restrict__ index, unsigned long long q, LayerPtrs layers, int n_layers,
                     uint32_t root, const unsigned long long* __restrict__ leaves, int S, unsigned long long width,
                     unsigned long long* __restrict__ out, uint32_t* __restrict__ out_of_range) {
  const unsigned long long j = (unsigned long long)blockIdx.x * DEC_THREADS + threadIdx.x;
  if (j >= q) return;
  unsigned long long idx = index[j];
  if (idx >= width) {
    *out_of_range = 1u;
    out[j] = 0ull;
    return;
  }
  uint32_t cur = root;
  for (int layer = n_layers - 1; layer >= 0; --layer) {
    const uint2 nd = __ldg(layers.nodes[layer] + (cur & IDX_MASK));
    const uint32_t m = (cur >> 29) & 1u, t = (cur >> 30) & 1u;
    const unsigned long long size = 1ull << layer;  // leaves under the logical left child (:272)
    uint32_t side = 0;
    if (idx >= size) {
      idx -= size;
      side = 1;
    }
    const uint32_t child = (side ^ m) ? nd.y : nd.x;
    cur = compose(child, m, t);
  }
  out[j] = apply_leaf(__ldg(leaves + (cur & IDX_MASK)), cur, S);
}

// width() for a tree that was not built here: nulls exist only on the logical right
// edge, so walk that edge: a present second child means the first subtree is full.
__global__ void width_kernel(LayerPtrs layers, int n_layers, uint32_t root, unsigned long long* __restrict__ out) {
  if (threadIdx.x || blockIdx.x) return;
  unsigned long long w = 0;
  uint32_t cur = root;
  if (ptr_is_null(cur)) {
    *out = 0;
    return;
  }
  for (int layer = n_layers - 1; layer >= 0; --layer) {
    const uint2 nd = layers.nodes[layer][cur & IDX_MASK];
    const uint32_t m = (cur >> 29) & 1u, t = (cur >> 30) & 1u;
    const uint32_t first = m ? nd.y : nd.x, second = m ? nd.x : nd.y;
    if (!ptr_is_null(second)) {
      w += 1ull << layer;
      cur = compose(second, m, t);
    } else {
      cur = compose(first, m, t);
    }
  }
  *out = w + 1;
}

// ---- host --------------------------------------------------------------------------

static int layer_ptrs(const Tree& t, LayerPtrs& lp) {
  if (t.layers.size() > MAX_LAYERS) return t.fail(STB_ERR_INVALID_ARG, "too many layers");
  for (size_t k = 0; k < t.layers.size(); ++k) lp.nodes[k] = t.layers[k].nodes.ptr;
  return STB_OK;
}

int compute_width(Tree& t) {
  LayerPtrs lp{};
  STB_TRY(layer_ptrs(t, lp));
  DevBuf<unsigned long long> d;
  STB_CUDA(t, d.alloc(1, t.stream));
  {
    Launch l(t, "width_walk");
    width_kernel<<<1, 32, 0, t.stream>>>(lp, (int)t.layers.size(), t.root, d.ptr);
  }
  unsigned long long w = 0;
  STB_CUDA(t, cudaMemcpyAsync(&w, d.ptr, 8, cudaMemcpyDeviceToHost, t.stream));
  STB_CUDA(t, cudaStreamSynchronize(t.stream));
  STB_CUDA(t, cudaGetLastError());
  t.width = w;
  return STB_OK;
}

int decode_range(const Tree& tc, uint64_t first, uint64_t count, unsigned long long* d_out, char* d_ascii) {
  Tree& t = const_cast<Tree&>(tc);
  if (!t.built) return t.fail(STB_ERR_NOT_BUILT, "tree is empty");
  if (count == 0) return STB_OK;
  if (first + count > t.width || first + count < first) return t.fail(STB_ERR_OUT_OF_RANGE, "decode range exceeds width()");
  cudaStream_t st = t.stream;
  const int L = (int)t.layers.size();
  const uint64_t last = first + count - 1;
  STB_CUDA(t, t.decode_a.ensure(count + 2, st));
  STB_CUDA(t, t.decode_b.ensure(count + 2, st));
  uint32_t* cur = t.decode_a.ptr;
  uint32_t* nxt = t.decode_b.ptr;
  STB_CUDA(t, cudaMemcpyAsync(cur, &t.root, 4, cudaMemcpyHostToDevice, st));
  uint64_t lo_cur = 0;
  // pointer level k refers to node layer k and covers 2^(k+1) leaves; level -1 = leaf pointers
  for (int k = L - 1; k >= 0; --k) {
    const uint64_t lo_next = first >> k, hi_next = last >> k;
    const uint64_t n_next = hi_next - lo_next + 1;
    {
      Launch l(t, "expand_level");
      expand_kernel<<<(unsigned)ceil_div(n_next, DEC_THREADS), DEC_THREADS, 0, st>>>(cur, lo_cur, t.layers[k].nodes.ptr, nxt, lo_next, n_next);
    }
    std::swap(cur, nxt);
    lo_cur = lo_next;
  }
  if (d_out) {
    Launch l(t, "leaves_out");
    leaves_out_kernel<<<(unsigned)ceil_div(count, DEC_THREADS), DEC_THREADS, 0, st>>>(cur, count, t.leaves.ptr, t.S, d_out);
  }
  if (d_ascii) {
    Launch l(t, "ascii_out");
    ascii_out_kernel<<<(unsigned)ceil_div(count, ASCII_TILE), DEC_THREADS, 0, st>>>(cur, count, t.leaves.ptr, t.S, d_ascii);
  }
  STB_CUDA(t, cudaStreamSynchronize(st));  // t.root was copied from host memory above
  STB_CUDA(t, cudaGetLastError());
  return STB_OK;
}

int random_access(const Tree& tc, const unsigned long long* d_index, uint64_t q, unsigned long long* d_out) {
  Tree& t = const_cast<Tree&>(tc);
  if (!t.built) return t.fail(STB_ERR_NOT_BUILT, "tree is empty");
  if (q == 0) return STB_OK;
  LayerPtrs lp{};
  STB_TRY(layer_ptrs(t, lp));
  DevBuf<uint32_t> flag;
  STB_CUDA(t, flag.alloc(1, t.stream));
  STB_CUDA(t, cudaMemsetAsync(flag.ptr, 0, 4, t.stream));
  {
    Launch l(t, "random_access");
    random_access_kernel<<<(unsigned)ceil_div(q, DEC_THREADS), DEC_THREADS, 0, t.stream>>>(
        d_index, q, lp, (int)t.layers.size(), t.root, t.leaves.ptr, t.S, t.width, d_out, flag.ptr);
  }
  uint32_t h = 0;
  STB_CUDA(t, cudaMemcpyAsync(&h, flag.ptr, 4, cudaMemcpyDeviceToHost, t.stream));
  STB_CUDA(t, cudaStreamSynchronize(t.stream));
  STB_CUDA(t, cudaGetLastError());
  if (h) return t.fail(STB_ERR_OUT_OF_RANGE, "random access index >= width()");
  return STB_OK;
}

}  // namespace stb

// ba_p2p.cu — the multi-GPU exchanges of the bundle adjuster as kernels over NVLink / NVSwitch peer memory.
//
// Per trial round the point-sharded adjuster sums the Schur-reduced camera systems of all candidates over
// the ranks, sums a handful of scalars, and once per outer iteration gathers the chi2 shards for the exact median.  The
// messages are small (a few MB at most, 80 bytes at least), so the cost of a library collective is its latency.  Here
// every exchange is ONE kernel per rank that writes straight into its peers' memory (cudaIpc-mapped, one process per
// GPU) in "LL" form: every double travels as the 16-byte line {lo, tag, hi, tag} whose two 8-byte halves validate
// themselves, so the receiver polls the data in its OWN memory and no flag, fence or second round trip is needed:
//
//   all-reduce (two-shot):  rank r pushes slice s of its contribution into inbox[r] of rank s; rank s adds the world
//                           contributions of its slice in rank order (one owner per slice: every rank ends up with the same
//                           bits, and they do not depend on timing) and pushes the sum into outbox[s] of every rank;
//   all-gather:             rank r pushes its shard into box[r] of every rank.
//
// The tag is the call number of the buffer (identical on all ranks: the call sequence is collective); two parities of
// every box alternate, which is enough because a rank can run at most one call ahead of its slowest peer.
#include "ba_types.cuh"

namespace mcp {

struct P2pPeers { uint4* base[8]; };           // the exchange buffer of every rank as mapped into this process

__device__ __forceinline__ void p2p_store(uint4* line, double v, unsigned tag)
{
  const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(line), "r"(lo), "r"(tag), "r"(hi), "r"(tag) : "memory");
}
__device__ __forceinline__ double p2p_wait(const uint4* line, unsigned tag)
{
  uint4 v;
  for (;;) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(line) : "memory");
    if (v.y == tag && v.w == tag) break;
  }
  return __hiloint2double((int)v.z, (int)v.x);
}

// lines per parity of the all-reduce boxes for `cnt` doubles: inbox world x slice, outbox world x slice
__host__ __device__ inline size_t p2p_slice(size_t cnt, int world) { return (cnt + world - 1) / world; }

__device__ __forceinline__ uint4 p2p_load(const uint4* line)
{
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(line) : "memory");
  return v;
}
// N lines (stride `step`) polled together: all loads of a round are in flight at once, only the missing ones are re-read
template <int N>
__device__ __forceinline__ void p2p_wait_n(const uint4* line, size_t step, int n, unsigned tag, double (&out)[N])
{
  uint4 v[N];
  unsigned pending = (n >= 32) ? 0xffffffffu : ((1u << n) - 1u);
  while (pending) {
#pragma unroll
    for (int k = 0; k < N; k++) if (pending & (1u << k)) v[k] = p2p_load(line + (size_t)k * step);
#pragma unroll
    for (int k = 0; k < N; k++) if ((pending & (1u << k)) && v[k].y == tag && v[k].w == tag) pending &= ~(1u << k);
  }
#pragma unroll
  for (int k = 0; k < N; k++) out[k] = (k < n) ? __hiloint2double((int)v[k].z, (int)v[k].x) : 0.0;
}

// buf[0..cnt) := sum over the ranks of buf[0..cnt), in place.  Layout of one parity: [inbox: world x slice | outbox: world x slice].
__global__ void __launch_bounds__(512, 2) k_p2p_allreduce(double* __restrict__ buf, size_t cnt, P2pPeers peers, size_t box_off, int rank, int world, unsigned tag)
{
  const size_t slice = p2p_slice(cnt, world);
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
  // 1. my contribution to every slice goes to the slice's owner
  for (size_t e = gtid; e < cnt; e += gsz) {
    const int s = (int)(e / slice);
    const size_t off = e - (size_t)s * slice;
    p2p_store(peers.base[s] + box_off + (size_t)rank * slice + off, buf[e], tag);
  }
  // 2. my slice: add the contributions in rank order, hand the sum to everybody
  const size_t lo = (size_t)rank * slice, hi = lo + slice < cnt ? lo + slice : cnt;
  const uint4* inbox = peers.base[rank] + box_off;
  for (size_t off = gtid; lo + off < hi; off += gsz) {
    double c[8];
    p2p_wait_n<8>(inbox + off, slice, world, tag, c);
    double sum = 0.0;
#pragma unroll
    for (int r = 0; r < 8; r++) if (r < world) sum += c[r];
    for (int p = 0; p < world; p++) p2p_store(peers.base[p] + box_off + (size_t)(world + rank) * slice + off, sum, tag);
  }
  // 3. collect the reduced slices (four lines per thread in flight)
  const uint4* outbox = peers.base[rank] + box_off + (size_t)world * slice;
  for (size_t e = gtid; e < cnt; e += 4 * gsz) {
    double c[4];
    const size_t left = (cnt - e + gsz - 1) / gsz;
    const int n = left < 4 ? (int)left : 4;
    p2p_wait_n<4>(outbox + e, gsz, n, tag, c);
#pragma unroll
    for (int k = 0; k < 4; k++) if (k < n) buf[e + (size_t)k * gsz] = c[k];
  }
}

// every rank's shard [bounds[r], bounds[r+1]) of the array ends up complete on every rank.  `which` < 0: the array is
// d.chi2[ctrl->cur] (the accepted state, known on the device only); stride = doubles per element.
struct P2pBounds { int b[9]; };
__global__ void __launch_bounds__(512, 2) k_p2p_allgather(BaDev d, double* __restrict__ arr_or_null, int stride, P2pBounds bounds, P2pPeers peers, size_t box_off,
                                                       int rank, int world, unsigned tag)
{
  double* arr = arr_or_null ? arr_or_null : d.chi2[d.ctrl->cur];
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsz = (size_t)gridDim.x * blockDim.x;
  const size_t lo = (size_t)bounds.b[rank] * stride, hi = (size_t)bounds.b[rank + 1] * stride;
  for (size_t e = lo + gtid; e < hi; e += gsz) {
    const double v = arr[e];
    for (int p = 0; p < world; p++) if (p != rank) p2p_store(peers.base[p] + box_off + e, v, tag);
  }
  const uint4* box = peers.base[rank] + box_off;
  const size_t n = (size_t)bounds.b[world] * stride;
  for (size_t e0 = gtid; e0 < n; e0 += 4 * gsz) {
    // four lines per thread in flight; own shard needs no wait
    uint4 v[4];
    unsigned pending = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { const size_t e = e0 + (size_t)k * gsz; if (e < n && (e < lo || e >= hi)) pending |= 1u << k; }
    const unsigned want = pending;
    while (pending) {
#pragma unroll
      for (int k = 0; k < 4; k++) if (pending & (1u << k)) v[k] = p2p_load(box + e0 + (size_t)k * gsz);
#pragma unroll
      for (int k = 0; k < 4; k++) if ((pending & (1u << k)) && v[k].y == tag && v[k].w == tag) pending &= ~(1u << k);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) if (want & (1u << k)) arr[e0 + (size_t)k * gsz] = __hiloint2double((int)v[k].z, (int)v[k].x);
  }
}

void launch_p2p_allreduce(double* buf, size_t cnt, const P2pPeers& peers, size_t box_off, int rank, int world, unsigned tag, cudaStream_t s)
{
  if (!cnt) return;
  int grid = (int)((cnt + 511) / 512);
  if (grid > 144) grid = 144;                     // ALL CTAs must be resident at once (they wait for their peers' CTAs): <= 1 per SM, 2 would fit
  k_p2p_allreduce<<<grid, 512, 0, s>>>(buf, cnt, peers, box_off, rank, world, tag);
}
void launch_p2p_allgather(const BaDev& d, double* arr_or_null, int stride, const int* bounds, const P2pPeers& peers, size_t box_off, int rank, int world,
                          unsigned tag, cudaStream_t s)
{
  P2pBounds b;
  for (int r = 0; r <= 8; r++) b.b[r] = bounds[r < world ? r : world];
  const size_t n = (size_t)bounds[world] * stride;
  if (!n) return;
  int grid = (int)((n / 4 + 511) / 512);
  if (grid > 144) grid = 144;
  if (grid < 1) grid = 1;
  k_p2p_allgather<<<grid, 512, 0, s>>>(d, arr_or_null, stride, b, peers, box_off, rank, world, tag);
}

}  // namespace mcp

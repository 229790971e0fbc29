// slab.cu -- z-slab plumbing for multi-GPU runs: halo pack / unpack for the
// width-2 z halos, and the block reshuffles on either side of the all-to-all transpose that
// brackets the z-direction GEMMs of the fast-diagonalisation solve.  The exchange itself is
// NCCL (torch.distributed) -- these kernels only make the strided columns contiguous.
#include "axb_common.cuh"

namespace {

// pack: left buffer  <- owned columns [ku0, ku0+w)      (what the LEFT neighbour needs)
//       right buffer <- owned columns [ku1-w, ku1)      (what the RIGHT neighbour needs)
__global__ void k_halo_pack(GridD g, const double* __restrict__ f, double* __restrict__ bl, double* __restrict__ br,
                            int w) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.nr * w) return;
  const int j = idx / w, c = idx % w;
  const double* row = f + (long long)j * g.ld;
  if (bl) bl[idx] = row[g.ku0 + c];
  if (br) br[idx] = row[g.ku1 - w + c];
}
// unpack: halo columns [ku0-w, ku0) <- left buffer - shift ; [ku1, ku1+w) <- right buffer + shift
__global__ void k_halo_unpack(GridD g, double* __restrict__ f, const double* __restrict__ bl,
                              const double* __restrict__ br, int w, double shift) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.nr * w) return;
  const int j = idx / w, c = idx % w;
  double* row = f + (long long)j * g.ld;
  if (bl) row[g.ku0 - w + c] = bl[idx] - shift;
  if (br) row[g.ku1 + c] = br[idx] + shift;
}

// halo exchange over peer memory: my first / last `w` owned columns go straight into the right halo of the
// left neighbour / the left halo of the right neighbour (all slabs share one layout)
__global__ void k_halo_put(GridD g, const double* __restrict__ f, double* __restrict__ left_peer,
                           double* __restrict__ right_peer, int w, double shift) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= g.nr * w) return;
  const int j = idx / w, c = idx % w;
  const long long row = (long long)j * g.ld;
  if (left_peer) left_peer[row + g.ku1 + c] = f[row + g.ku0 + c] + shift;
  if (right_peer) right_peer[row + g.ku0 - w + c] = f[row + g.ku1 - w + c] - shift;
}

// slab (nr x nzl, pitch ld)  ->  P contiguous blocks, block q = rows [q*nrl, (q+1)*nrl) x nzl
__global__ void k_slab_to_blocks(int nr, int nzl, long long ld, const double* __restrict__ slab,
                                 double* __restrict__ blocks) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (k < nzl) blocks[(long long)j * nzl + k] = slab[(long long)j * ld + k];
}
// P received blocks (block q = my rows x columns of rank q, nrl x nzl) -> rows (nrl x P*nzl)
__global__ void k_blocks_to_rows(int nrl, int nzl, int P, const double* __restrict__ blocks, double* __restrict__ rows,
                                 long long ldr) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y, q = blockIdx.z;
  if (k < nzl) rows[(long long)j * ldr + (long long)q * nzl + k] = blocks[((long long)q * nrl + j) * nzl + k];
}
__global__ void k_rows_to_blocks(int nrl, int nzl, int P, const double* __restrict__ rows, long long ldr,
                                 double* __restrict__ blocks) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y, q = blockIdx.z;
  if (k < nzl) blocks[((long long)q * nrl + j) * nzl + k] = rows[(long long)j * ldr + (long long)q * nzl + k];
}
__global__ void k_blocks_to_slab(int nr, int nzl, long long ld, const double* __restrict__ blocks,
                                 double* __restrict__ slab) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (k < nzl) slab[(long long)j * ld + k] = blocks[(long long)j * nzl + k];
}

}  // namespace

// ---- transposes over peer memory (NVLink): every rank stores its blocks straight into the peers'
// destination buffers in their final layout, so the all-to-all needs no pack / unpack passes and no
// library call.  peers.p[q] is rank q's destination buffer (mapped into this process).
struct PeerPtrs {
  double* p[AXB_MAX_PEERS];
};
__global__ void __launch_bounds__(256)
    k_peer_block_put(PeerPtrs peers, int P, int me, long long dst_off, long long ld_dst, const double* __restrict__ src,
                     long long src_peer_stride, long long ld_src, int rows, int cols, bool vec) {
  const int q = (me + 1 + blockIdx.x) % P;                    // peers interleaved: all links busy at once
  const int i = blockIdx.y;
  const double* s = src + (long long)q * src_peer_stride + (long long)i * ld_src;
  double* d = peers.p[q] + dst_off + (long long)i * ld_dst;
  const int c0 = (blockIdx.z * 256 + threadIdx.x) * 4;
  if (c0 >= cols) return;
  if (vec && c0 + 3 < cols) {
    const double2 a = *reinterpret_cast<const double2*>(s + c0);
    const double2 b = *reinterpret_cast<const double2*>(s + c0 + 2);
    *reinterpret_cast<double2*>(d + c0) = a;
    *reinterpret_cast<double2*>(d + c0 + 2) = b;
  } else {
    for (int c = c0; c < cols && c < c0 + 4; ++c) d[c] = s[c];
  }
}

// ---- r-slab (row) halos over peer memory: the first / last `w` owned rows of up to 8 fields go straight into the
// upper halo rows of the lower neighbour / the lower halo rows of the upper neighbour.  Every rank stores its block
// as (halo + nrl + halo) rows of pitch ld, so the rows are contiguous runs on both sides.
struct RowHaloFields {
  const double* src[8];
  double* lower[8];
  double* upper[8];
};
template <bool GET>
__global__ void __launch_bounds__(256)
    k_row_halo_put(RowHaloFields f, long long ld, int nz, int nrl, int halo, int w, bool vec) {
  const int fi = blockIdx.z, up = blockIdx.y;
  double* peer = up ? f.upper[fi] : f.lower[fi];
  if (!peer) return;
  double* mine = const_cast<double*>(f.src[fi]);
  // put: my edge rows -> the neighbour's halo rows;  get: the neighbour's edge rows -> my halo rows
  const double* s = GET ? peer + (long long)(up ? halo : halo + nrl - w) * ld
                        : mine + (long long)(up ? halo + nrl - w : halo) * ld;
  double* d = GET ? mine + (long long)(up ? halo + nrl : halo - w) * ld
                  : peer + (long long)(up ? halo - w : halo + nrl) * ld;
  const int per_row = (nz + 1) / 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < w * per_row; i += gridDim.x * blockDim.x) {
    const int r = i / per_row, c = 2 * (i - r * per_row);
    const long long o = (long long)r * ld + c;
    if (vec && c + 1 < nz) {
      *reinterpret_cast<double2*>(d + o) = *reinterpret_cast<const double2*>(s + o);
    } else {
      d[o] = s[o];
      if (c + 1 < nz) d[o + 1] = s[o + 1];
    }
  }
}

// ---- rank synchronisation by flags in peer-mapped memory (no collective library, CUDA-graph capturable) -----------
// Every rank owns a control block (AXB_CTL_WORDS 64-bit words, zero-initialised, mapped into all processes):
//   [0] flag written by the LOWER neighbour, [1] flag written by the UPPER neighbour   (row-halo exchange)
//   [8, 8 + 16)        barrier flags, one per rank
//   [32, 32 + 2*16*2)  all-reduce slots: [parity][rank] x {value, epoch}
// Epochs live in a LOCAL device counter block (3 words + tickets) so that a captured graph advances them itself.
// Spins give up after ~2 s of clock64 and raise word [7] of the local counters instead of hanging the GPU.
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool spin_until(const unsigned long long* flag, unsigned long long e, unsigned long long* err) {
  const long long t0 = clock64();
  while (ld_acquire_sys(flag) < e) {
    if (clock64() - t0 > 4000000000LL) {
      atomicExch(err, 1ULL);
      return false;
    }
    __nanosleep(64);
  }
  return true;
}

struct CtlPtrs {
  unsigned long long* p[AXB_MAX_PEERS];
};

// all-rank barrier: one block, thread q talks to rank q
__global__ void k_peer_sync(CtlPtrs ctl, int P, int me, unsigned long long* cnt) {
  __shared__ unsigned long long e_s;
  if (threadIdx.x == 0) e_s = cnt[1] + 1;
  __syncthreads();
  const unsigned long long e = e_s;
  const int q = threadIdx.x;
  if (q < P) {
    __threadfence_system();
    st_release_sys(ctl.p[q] + 8 + me, e);
    spin_until(ctl.p[me] + 8 + q, e, cnt + 7);
  }
  __syncthreads();
  if (threadIdx.x == 0) cnt[1] = e;
}

// MAX all-reduce of one double (in place): thread q sends to / receives from rank q
__global__ void k_peer_allreduce_max(CtlPtrs ctl, int P, int me, double* val, unsigned long long* cnt) {
  __shared__ unsigned long long e_s;
  __shared__ double v_s[AXB_MAX_PEERS];
  if (threadIdx.x == 0) e_s = cnt[2] + 1;
  __syncthreads();
  const unsigned long long e = e_s;
  const int q = threadIdx.x;
  if (q < P) {
    const double mine = *val;
    unsigned long long* dst = ctl.p[q] + 32 + ((e & 1) * AXB_MAX_PEERS + me) * 2;
    dst[0] = (unsigned long long)__double_as_longlong(mine);
    __threadfence_system();
    st_release_sys(dst + 1, e);
    unsigned long long* src = ctl.p[me] + 32 + ((e & 1) * AXB_MAX_PEERS + q) * 2;
    const bool ok = spin_until(src + 1, e, cnt + 7);
    v_s[q] = ok ? __longlong_as_double((long long)*reinterpret_cast<volatile unsigned long long*>(src)) : mine;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = v_s[0];
    for (int i = 1; i < P; ++i) m = fmax(m, v_s[i]);
    *val = m;
    cnt[2] = e;
  }
}

// row-halo exchange in one launch: block (0,0,0) tells both neighbours "my rows of epoch e are ready", every block
// waits for the neighbour it reads from and then pulls that neighbour's edge rows into this rank's halo rows.
__global__ void __launch_bounds__(256)
    k_row_halo_exchange(RowHaloFields f, long long ld, int nz, int nrl, int halo, int w, bool vec,
                        unsigned long long* ctl_mine, unsigned long long* ctl_lower, unsigned long long* ctl_upper,
                        unsigned long long* cnt) {
  __shared__ unsigned long long e_s;
  __shared__ int ok_s;
  const int fi = blockIdx.z, up = blockIdx.y;
  const int nblocks = gridDim.x * gridDim.y * gridDim.z;
  if (threadIdx.x == 0) {
    const unsigned long long e = cnt[0] + 1;          // bumped by the last block to leave (ticket in cnt[4])
    e_s = e;
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
      __threadfence_system();
      if (ctl_lower) st_release_sys(ctl_lower + 1, e);  // I am the lower neighbour's UPPER neighbour
      if (ctl_upper) st_release_sys(ctl_upper + 0, e);
    }
    const unsigned long long* peer_ctl = up ? ctl_upper : ctl_lower;
    ok_s = peer_ctl ? (spin_until(ctl_mine + (up ? 1 : 0), e, cnt + 7) ? 1 : 0) : 0;
  }
  __syncthreads();
  double* peer = up ? f.upper[fi] : f.lower[fi];
  if (peer && ok_s) {
    double* mine = const_cast<double*>(f.src[fi]);
    const double* s = peer + (long long)(up ? halo : halo + nrl - w) * ld;
    double* d = mine + (long long)(up ? halo + nrl : halo - w) * ld;
    const int per_row = (nz + 1) / 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < w * per_row; i += gridDim.x * blockDim.x) {
      const int r = i / per_row, c = 2 * (i - r * per_row);
      const long long o = (long long)r * ld + c;
      if (vec && c + 1 < nz) {
        *reinterpret_cast<double2*>(d + o) = *reinterpret_cast<const double2*>(s + o);
      } else {
        d[o] = s[o];
        if (c + 1 < nz) d[o + 1] = s[o + 1];
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long t = atomicAdd(cnt + 4, 1ULL);
    if (t == (unsigned long long)nblocks - 1) {
      cnt[4] = 0;
      cnt[0] = e_s;
      __threadfence();
    }
  }
}

extern "C" {

int axb_peer_sync(int P, int me, const uint64_t* ctl_ptrs, uint64_t* counters, axb_stream_t s) {
  if (P < 1 || P > AXB_MAX_PEERS || me < 0 || me >= P || !ctl_ptrs || !counters) return AXB_EINVAL;
  CtlPtrs c;
  for (int q = 0; q < AXB_MAX_PEERS; ++q) c.p[q] = q < P ? reinterpret_cast<unsigned long long*>(ctl_ptrs[q]) : nullptr;
  for (int q = 0; q < P; ++q)
    if (!c.p[q]) return AXB_EINVAL;
  k_peer_sync<<<1, 32, 0, (cudaStream_t)s>>>(c, P, me, reinterpret_cast<unsigned long long*>(counters));
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_peer_allreduce_max(int P, int me, const uint64_t* ctl_ptrs, double* value, uint64_t* counters, axb_stream_t s) {
  if (P < 1 || P > AXB_MAX_PEERS || me < 0 || me >= P || !ctl_ptrs || !counters || !value) return AXB_EINVAL;
  CtlPtrs c;
  for (int q = 0; q < AXB_MAX_PEERS; ++q) c.p[q] = q < P ? reinterpret_cast<unsigned long long*>(ctl_ptrs[q]) : nullptr;
  for (int q = 0; q < P; ++q)
    if (!c.p[q]) return AXB_EINVAL;
  k_peer_allreduce_max<<<1, 32, 0, (cudaStream_t)s>>>(c, P, me, value, reinterpret_cast<unsigned long long*>(counters));
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_row_halo_exchange(int nfields, const uint64_t* mine, const uint64_t* lower_peer, const uint64_t* upper_peer,
                          int64_t ld, int nz, int nrl, int halo, int width, uint64_t* ctl_mine, uint64_t* ctl_lower,
                          uint64_t* ctl_upper, uint64_t* counters, axb_stream_t s) {
  if (nfields < 1 || nfields > 8 || !mine || !lower_peer || !upper_peer || nz < 1 || ld < nz || width < 1 ||
      width > halo || nrl < width || !ctl_mine || !counters)
    return AXB_EINVAL;
  RowHaloFields f;
  bool vec = (ld % 2 == 0);
  for (int i = 0; i < 8; ++i) {
    f.src[i] = nullptr; f.lower[i] = nullptr; f.upper[i] = nullptr;
    if (i >= nfields) continue;
    f.src[i] = reinterpret_cast<const double*>(mine[i]);
    f.lower[i] = reinterpret_cast<double*>(lower_peer[i]);
    f.upper[i] = reinterpret_cast<double*>(upper_peer[i]);
    if (!f.src[i]) return AXB_EINVAL;
    if ((f.lower[i] != nullptr) != (ctl_lower != nullptr) || (f.upper[i] != nullptr) != (ctl_upper != nullptr))
      return AXB_EINVAL;
    vec = vec && axb_al16(f.src[i]) && axb_al16(f.lower[i]) && axb_al16(f.upper[i]);
  }
  const int work = width * ((nz + 1) / 2);
  int bx = (work + 255) / 256;
  if (bx > 32) bx = 32;
  k_row_halo_exchange<<<dim3(bx, 2, nfields), 256, 0, (cudaStream_t)s>>>(
      f, ld, nz, nrl, halo, width, vec, reinterpret_cast<unsigned long long*>(ctl_mine),
      reinterpret_cast<unsigned long long*>(ctl_lower), reinterpret_cast<unsigned long long*>(ctl_upper),
      reinterpret_cast<unsigned long long*>(counters));
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

static int row_halo(bool get, int nfields, const uint64_t* src, const uint64_t* lower_peer, const uint64_t* upper_peer,
                    int64_t ld, int nz, int nrl, int halo, int width, axb_stream_t s) {
  if (nfields < 1 || nfields > 8 || !src || !lower_peer || !upper_peer || nz < 1 || ld < nz || width < 1 || width > halo ||
      nrl < width)
    return AXB_EINVAL;
  RowHaloFields f;
  bool vec = (ld % 2 == 0), any = false;
  for (int i = 0; i < 8; ++i) {
    f.src[i] = nullptr; f.lower[i] = nullptr; f.upper[i] = nullptr;
    if (i >= nfields) continue;
    f.src[i] = reinterpret_cast<const double*>(src[i]);
    f.lower[i] = reinterpret_cast<double*>(lower_peer[i]);
    f.upper[i] = reinterpret_cast<double*>(upper_peer[i]);
    if (!f.src[i]) return AXB_EINVAL;
    vec = vec && axb_al16(f.src[i]) && axb_al16(f.lower[i]) && axb_al16(f.upper[i]);
    any = any || f.lower[i] || f.upper[i];
  }
  if (!any) return AXB_OK;
  const int work = width * ((nz + 1) / 2);
  int bx = (work + 255) / 256;
  if (bx > 64) bx = 64;
  if (get) k_row_halo_put<true><<<dim3(bx, 2, nfields), 256, 0, (cudaStream_t)s>>>(f, ld, nz, nrl, halo, width, vec);
  else k_row_halo_put<false><<<dim3(bx, 2, nfields), 256, 0, (cudaStream_t)s>>>(f, ld, nz, nrl, halo, width, vec);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}
int axb_row_halo_put(int nfields, const uint64_t* src, const uint64_t* lower_peer, const uint64_t* upper_peer, int64_t ld,
                     int nz, int nrl, int halo, int width, axb_stream_t s) {
  return row_halo(false, nfields, src, lower_peer, upper_peer, ld, nz, nrl, halo, width, s);
}
int axb_row_halo_get(int nfields, const uint64_t* mine, const uint64_t* lower_peer, const uint64_t* upper_peer, int64_t ld,
                     int nz, int nrl, int halo, int width, axb_stream_t s) {
  return row_halo(true, nfields, mine, lower_peer, upper_peer, ld, nz, nrl, halo, width, s);
}

int axb_halo_pack(const axb_grid_t* g, const double* f, double* buf_left, double* buf_right, int width,
                  axb_stream_t s) {
  if (!f || width < 1) return AXB_EINVAL;
  int rc = axb_check_grid(g);
  if (rc) return rc;
  const GridD d = to_dev(g);
  if (d.ku1 - d.ku0 < width) return AXB_EINVAL;
  const int n = d.nr * width;
  k_halo_pack<<<(n + 255) / 256, 256, 0, s>>>(d, f, buf_left, buf_right, width);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_halo_unpack(const axb_grid_t* g, double* f, const double* buf_left, const double* buf_right,
                    int width, double shift, axb_stream_t s) {
  if (!f || width < 1) return AXB_EINVAL;
  int rc = axb_check_grid(g);
  if (rc) return rc;
  const GridD d = to_dev(g);
  if ((buf_left && d.ku0 < width) || (buf_right && d.ku1 + width > d.nz)) return AXB_EINVAL;
  const int n = d.nr * width;
  k_halo_unpack<<<(n + 255) / 256, 256, 0, s>>>(d, f, buf_left, buf_right, width, shift);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_halo_put(const axb_grid_t* g, const double* f, double* left_peer_field, double* right_peer_field, int width,
                 double shift, axb_stream_t s) {
  if (!f || width < 1) return AXB_EINVAL;
  int rc = axb_check_grid(g);
  if (rc) return rc;
  const GridD d = to_dev(g);
  if (d.ku0 < width || d.ku1 + width > d.nz || d.ku1 - d.ku0 < width) return AXB_EINVAL;
  const int n = d.nr * width;
  k_halo_put<<<(n + 255) / 256, 256, 0, s>>>(d, f, left_peer_field, right_peer_field, width, shift);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_slab_to_blocks(int nr, int nzl, int64_t ld, int P, const double* slab, double* blocks,
                       axb_stream_t s) {
  if (!slab || !blocks || nr < 1 || nzl < 1 || P < 1 || nr % P) return AXB_EINVAL;
  k_slab_to_blocks<<<dim3((nzl + 127) / 128, nr), 128, 0, s>>>(nr, nzl, ld, slab, blocks);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}
int axb_blocks_to_rows(int nrl, int nzl, int P, const double* blocks, double* rows, int64_t ld_rows,
                       axb_stream_t s) {
  if (!rows || !blocks || nrl < 1 || nzl < 1 || P < 1) return AXB_EINVAL;
  k_blocks_to_rows<<<dim3((nzl + 127) / 128, nrl, P), 128, 0, s>>>(nrl, nzl, P, blocks, rows, ld_rows);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}
int axb_rows_to_blocks(int nrl, int nzl, int P, const double* rows, int64_t ld_rows, double* blocks,
                       axb_stream_t s) {
  if (!rows || !blocks || nrl < 1 || nzl < 1 || P < 1) return AXB_EINVAL;
  k_rows_to_blocks<<<dim3((nzl + 127) / 128, nrl, P), 128, 0, s>>>(nrl, nzl, P, rows, ld_rows, blocks);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}
int axb_blocks_to_slab(int nr, int nzl, int64_t ld, int P, const double* blocks, double* slab,
                       axb_stream_t s) {
  if (!slab || !blocks || nr < 1 || nzl < 1 || P < 1 || nr % P) return AXB_EINVAL;
  k_blocks_to_slab<<<dim3((nzl + 127) / 128, nr), 128, 0, s>>>(nr, nzl, ld, blocks, slab);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}

int axb_peer_block_put(int P, int me, const uint64_t* peer_ptrs, int64_t dst_off, int64_t ld_dst, const double* src,
                       int64_t src_peer_stride, int64_t ld_src, int rows, int cols, axb_stream_t s) {
  if (P < 1 || P > AXB_MAX_PEERS || me < 0 || me >= P || !peer_ptrs || !src || rows < 1 || cols < 1) return AXB_EINVAL;
  PeerPtrs pp;
  bool vec = axb_al16(src) && (ld_src % 2 == 0) && (ld_dst % 2 == 0) && (dst_off % 2 == 0) && (src_peer_stride % 2 == 0);
  for (int q = 0; q < P; ++q) {
    pp.p[q] = reinterpret_cast<double*>(peer_ptrs[q]);
    if (!pp.p[q]) return AXB_EINVAL;
    vec = vec && axb_al16(pp.p[q]);
  }
  k_peer_block_put<<<dim3(P, rows, (cols + 1023) / 1024), 256, 0, (cudaStream_t)s>>>(pp, P, me, dst_off, ld_dst, src,
                                                                                    src_peer_stride, ld_src, rows, cols, vec);
  AXB_LAUNCHED();
  AXB_RETURN_LAST();
}
}  // extern "C"

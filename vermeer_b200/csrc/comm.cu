// Multi-GPU frame exchange behind the C ABI: one process per GPU, every context renders the 32x32 tiles it
// owns (core/render.go:196-199 tiles dealt round-robin, owned_pixels() below), and ONE exchange at frame end brings the owned
// pixels to rank 0:
//
//   every rank   k_pack_owned      owned pixels of the row-major frame -> one dense send buffer (tile-major, 12 B per pixel)
//   rank r > 0   ncclSend -> 0     nown_r * 3 floats over NVLink / NVSwitch      } one ncclGroup on the context's stream
//   rank 0       ncclRecv <- r     into a staging buffer, rank after rank         }
//   rank 0       k_scatter_owned   staging -> row-major frame (ownership is disjoint: plain stores, no reduction)
//   rank 0       D2H of the complete frame into the caller's buffer (DMA straight into it when it is page-locked)
//
// Only owned pixels travel ((N-1)/N of the frame in total, all of it into rank 0), nothing is summed, so the gathered frame is
// bit-identical to a single-GPU render of the same iterations. The reference has no counterpart (one process, goroutines
// writing disjoint framebuffer ranges, core/render.go:127-129).
//
// NCCL is bound at run time with dlopen("libnccl.so.2") the first time a communicator is needed: the library keeps loading on
// hosts without NCCL (single-GPU use), and inside a process that already holds an NCCL (e.g. PyTorch's bundled one) the same
// copy is shared instead of a second one being loaded. Only the handful of entry points below are used; their prototypes and
// the two constants are restated here so that nccl.h is not needed to build.
#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "context.h"

namespace vg {

// ---- tile ownership -------------------------------------------------------------------------------------------------
// The reference's 32x32 tiles (render.go:196-199), tile (tx,ty) -> rank (tx + ty*k) % world with k >= tilesX the smallest
// stride coprime with `world`: a row-major deal whose row-to-row skew never aliases whole tile columns onto one rank, for any
// world size (k odd alone guarantees that only when gcd(k, world) = 1).
static int gcd_i(int a, int b) { return b == 0 ? a : gcd_i(b, a % b); }
int partition_stride(int tilesX, int world) {
  int k = tilesX;
  while (gcd_i(k, world) != 1) k++;
  return k;
}

void owned_pixels(int W, int H, int rank, int world, bool pixel_block, std::vector<int>& pix) {
  pix.clear();
  const int tilesX = (W + 31) / 32, tilesY = (H + 31) / 32;
  const int k = partition_stride(tilesX, world);
  for (int ty = 0; ty < tilesY; ty++)
    for (int tx = 0; tx < tilesX; tx++) {
      if ((tx + ty * k) % world != rank) continue;
      // Path order inside a tile decides which camera rays share a warp (render.cu: path_index). 8x4 pixel blocks, Morton
      // order inside a block (bits of l, low to high: x0 y0 x1 y1 x2): 2, 4, 8, 16 consecutive pixels are 2x1, 2x2, 4x2, 4x4
      // sub-blocks. Results are per pixel and do not depend on this order.
      if (pixel_block) {
        for (int b = 0; b < 32; b++)
          for (int l = 0; l < 32; l++) {
            const int lx = (l & 1) | ((l >> 1) & 2) | ((l >> 2) & 4), ly = ((l >> 1) & 1) | ((l >> 2) & 2);
            const int x = tx * 32 + (b & 3) * 8 + lx, y = ty * 32 + (b >> 2) * 4 + ly;
            if (x < W && y < H) pix.push_back(x + y * W);
          }
      } else {
        for (int j = 0; j < 32; j++)
          for (int i = 0; i < 32; i++) {
            const int x = tx * 32 + i, y = ty * 32 + j;
            if (x < W && y < H) pix.push_back(x + y * W);
          }
      }
    }
}

// ---- NCCL, bound at run time ----------------------------------------------------------------------------------------
struct NcclUniqueId { char internal[VG_COMM_ID_BYTES]; };  // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef struct ncclComm* NcclComm;
static const int kNcclFloat32 = 7;  // ncclFloat32
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  std::string error;
};
static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static bool nccl_load(std::string* err) {
  std::lock_guard<std::mutex> lock(g_nccl_mu);
  if (g_nccl.lib) return true;
  const char* override_path = std::getenv("VG_NCCL_LIB");
  const char* names[] = {override_path, "libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* n : names) {
    if (!n || !*n) continue;
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) {
    *err = std::string("NCCL is not available (dlopen libnccl.so.2: ") + (dlerror() ? dlerror() : "not found") + "); set VG_NCCL_LIB";
    return false;
  }
  NcclApi a;
  a.lib = lib;
#define VG_SYM(field, name)                                              \
  *reinterpret_cast<void**>(&a.field) = dlsym(lib, name);                \
  if (!a.field) {                                                        \
    *err = std::string("libnccl lacks ") + name;                         \
    dlclose(lib);                                                        \
    return false;                                                        \
  }
  VG_SYM(GetUniqueId, "ncclGetUniqueId")
  VG_SYM(CommInitRank, "ncclCommInitRank")
  VG_SYM(CommDestroy, "ncclCommDestroy")
  VG_SYM(Send, "ncclSend")
  VG_SYM(Recv, "ncclRecv")
  VG_SYM(GroupStart, "ncclGroupStart")
  VG_SYM(GroupEnd, "ncclGroupEnd")
  VG_SYM(GetErrorString, "ncclGetErrorString")
  VG_SYM(GetVersion, "ncclGetVersion")
#undef VG_SYM
  g_nccl = a;
  return true;
}

struct CommState {
  NcclComm comm = nullptr;
  int rank = 0, world = 1;
  // layout of the gathered staging buffer on rank 0, valid for (W, H, pixel_block) below
  int W = 0, H = 0, pixel_block = -1;
  std::vector<int> count;   // owned pixels per rank
  std::vector<int> offset;  // first staging pixel of rank r (ranks 1.., rank 0's pixels are already in place)
  std::vector<std::vector<int>> row_start;  // [rank][tilesY + 1] first owned-list index of every tile row (slices of the frame)
  DevBuf<int> pix_own;      // this rank's pixel list
  DevBuf<int> pix_others;   // rank 0: concatenated lists of ranks 1..N-1, in staging order
  DevBuf<float> send;       // packed owned pixels
  DevBuf<float> stage;      // rank 0: received pixels
  int n_others = 0;
  uint64_t gathers = 0;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  double last_ms = 0;
};

__global__ void k_pack_owned(const float* __restrict__ fb, const int* __restrict__ pix, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 3) return;
  out[i] = fb[(size_t)pix[i / 3] * 3 + i % 3];
}
__global__ void k_scatter_owned(const float* __restrict__ in, const int* __restrict__ pix, int n, float* __restrict__ fb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 3) return;
  fb[(size_t)pix[i / 3] * 3 + i % 3] = in[i];
}

void comm_destroy(vg_ctx* ctx) {
  CommState* cs = ctx->comm;
  if (!cs) return;
  if (cs->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(cs->comm);
  cs->pix_own.release(); cs->pix_others.release(); cs->send.release(); cs->stage.release();
  if (cs->e0) cudaEventDestroy(cs->e0);
  if (cs->e1) cudaEventDestroy(cs->e1);
  delete cs;
  ctx->comm = nullptr;
}

}  // namespace vg

using namespace vg;

#define CCUDA(call)                                          \
  do {                                                       \
    cudaError_t e_ = (call);                                 \
    if (e_ != cudaSuccess) return ctx->cuda_fail(e_, #call); \
  } while (0)
#define CNCCL(call)                                                                                   \
  do {                                                                                                \
    int r_ = (call);                                                                                  \
    if (r_ != 0) return ctx->fail(VG_ERR_COMM, std::string(#call) + ": " + g_nccl.GetErrorString(r_)); \
  } while (0)

extern "C" int vg_owned_pixels(int xres, int yres, int rank, int world, int pixel_block, int32_t* pix_out, int64_t cap) {
  if (xres <= 0 || yres <= 0 || world <= 0 || rank < 0 || rank >= world) return VG_ERR_INVALID;
  std::vector<int> pix;
  owned_pixels(xres, yres, rank, world, pixel_block != 0, pix);
  if (pix_out) {
    if ((int64_t)pix.size() > cap) return VG_ERR_INVALID;
    std::memcpy(pix_out, pix.data(), pix.size() * sizeof(int));
  }
  return (int)pix.size();
}

extern "C" int vg_comm_unique_id(vg_ctx* ctx, void* id_out) {
  if (!ctx || !id_out) return VG_ERR_INVALID;
  std::lock_guard<std::mutex> lock(ctx->mu);
  std::string err;
  if (!nccl_load(&err)) return ctx->fail(VG_ERR_COMM, err);
  NcclUniqueId id;
  CNCCL(g_nccl.GetUniqueId(&id));
  std::memcpy(id_out, &id, sizeof(id));
  return VG_OK;
}

extern "C" int vg_comm_init(vg_ctx* ctx, int rank, int world, const void* id) {
  if (!ctx) return VG_ERR_INVALID;
  std::lock_guard<std::mutex> lock(ctx->mu);
  if (world <= 0 || rank < 0 || rank >= world || (world > 1 && !id)) return ctx->fail(VG_ERR_INVALID, "vg_comm_init: bad rank/world/id");
  comm_destroy(ctx);
  CCUDA(cudaSetDevice(ctx->device));
  CommState* cs = new CommState();
  cs->rank = rank;
  cs->world = world;
  ctx->comm = cs;
  if (world > 1) {
    std::string err;
    if (!nccl_load(&err)) return ctx->fail(VG_ERR_COMM, err);
    NcclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    CNCCL(g_nccl.CommInitRank(&cs->comm, world, uid, rank));
  }
  cudaEventCreate(&cs->e0);
  cudaEventCreate(&cs->e1);
  // the image partition follows the communicator (vg_set_partition)
  ctx->rank = rank;
  ctx->world = world;
  render_invalidate(ctx);
  return VG_OK;
}

extern "C" int vg_comm_destroy(vg_ctx* ctx) {
  if (!ctx) return VG_ERR_INVALID;
  std::lock_guard<std::mutex> lock(ctx->mu);
  comm_destroy(ctx);
  return VG_OK;
}

extern "C" int vg_nccl_version(void) {
  std::string err;
  if (!nccl_load(&err)) return 0;
  int v = 0;
  g_nccl.GetVersion(&v);
  return v;
}

// (re)build the staging layout when the frame or the pixel order changed
static int comm_layout(vg_ctx* ctx, CommState* cs) {
  const int W = ctx->xres, H = ctx->yres, pb = ctx->opt_pixel_block ? 1 : 0;
  if (cs->W == W && cs->H == H && cs->pixel_block == pb) return VG_OK;
  std::vector<int> pix, others;
  cs->count.assign((size_t)cs->world, 0);
  cs->offset.assign((size_t)cs->world, 0);
  cs->row_start.assign((size_t)cs->world, std::vector<int>());
  for (int r = 0; r < cs->world; r++) {
    owned_pixels(W, H, r, cs->world, pb != 0, pix);
    cs->count[(size_t)r] = (int)pix.size();
    {
      const int tilesY = (H + 31) / 32;
      std::vector<int>& rsr = cs->row_start[(size_t)r];
      rsr.assign((size_t)tilesY + 1, (int)pix.size());
      for (int i = (int)pix.size() - 1; i >= 0; i--) rsr[(size_t)((pix[(size_t)i] / W) / 32)] = i;
      for (int ty = tilesY - 1; ty >= 0; ty--) rsr[(size_t)ty] = std::min(rsr[(size_t)ty], rsr[(size_t)ty + 1]);
    }
    if (r == cs->rank) {
      CCUDA(cs->pix_own.reserve(pix.size()));
      if (!pix.empty()) CCUDA(cudaMemcpyAsync(cs->pix_own.p, pix.data(), pix.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
      CCUDA(cs->send.reserve(pix.size() * 3));
      CCUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (cs->rank == 0 && r > 0) {
      cs->offset[(size_t)r] = (int)others.size();
      others.insert(others.end(), pix.begin(), pix.end());
    }
  }
  cs->n_others = (int)others.size();
  if (cs->rank == 0 && cs->n_others > 0) {
    CCUDA(cs->pix_others.reserve(others.size()));
    CCUDA(cudaMemcpyAsync(cs->pix_others.p, others.data(), others.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    CCUDA(cs->stage.reserve(others.size() * 3));
    CCUDA(cudaStreamSynchronize(ctx->stream));
  }
  cs->W = W; cs->H = H; cs->pixel_block = pb;
  return VG_OK;
}

static int comm_layout(vg_ctx* ctx, vg::CommState* cs);
namespace vg {
// The exchange for the tile rows [ty0, ty1) on `st` (the whole frame: 0 .. tilesY): pack, send / receive, scatter, and on rank 0 the
// D2H of those image rows into fb_out. `sync`: wait for it. The caller holds the context lock.
int comm_gather_rows(vg_ctx* ctx, cudaStream_t st, int ty0, int ty1, float* fb_out, bool sync) {
  CommState* cs = ctx->comm;
  if (!cs) return ctx->fail(VG_ERR_INVALID, "vg_gather_frame: no communicator (vg_comm_init)");
  if (cs->rank != ctx->rank || cs->world != ctx->world) return ctx->fail(VG_ERR_INVALID, "vg_gather_frame: vg_set_partition changed the partition after vg_comm_init");
  float* fb = nullptr;
  int rc = render_fb_device(ctx, &fb);
  if (rc != VG_OK) return rc;
  if (cs->world > 1) {
    rc = comm_layout(ctx, cs);
    if (rc != VG_OK) return rc;
    const std::vector<int>& mine = cs->row_start[(size_t)cs->rank];
    const int o0 = mine[(size_t)ty0], nown = mine[(size_t)ty1] - mine[(size_t)ty0];
    if (cs->rank != 0) {
      if (nown > 0) {
        k_pack_owned<<<(nown * 3 + 255) / 256, 256, 0, st>>>(fb, cs->pix_own.p + o0, nown, cs->send.p + (size_t)o0 * 3);
        CCUDA(cudaGetLastError());
        CNCCL(g_nccl.Send(cs->send.p + (size_t)o0 * 3, (size_t)nown * 3, kNcclFloat32, 0, cs->comm, st));
      }
    } else {
      CNCCL(g_nccl.GroupStart());
      for (int r = 1; r < cs->world; r++) {
        const std::vector<int>& rr = cs->row_start[(size_t)r];
        const int q0 = rr[(size_t)ty0], n = rr[(size_t)ty1] - rr[(size_t)ty0];
        if (n <= 0) continue;
        const int e = g_nccl.Recv(cs->stage.p + ((size_t)cs->offset[(size_t)r] + (size_t)q0) * 3, (size_t)n * 3, kNcclFloat32, r, cs->comm, st);
        if (e != 0) {
          g_nccl.GroupEnd();
          return ctx->fail(VG_ERR_COMM, std::string("ncclRecv: ") + g_nccl.GetErrorString(e));
        }
      }
      CNCCL(g_nccl.GroupEnd());
      for (int r = 1; r < cs->world; r++) {  // one scatter per rank: that rank's slice is one run of the staging buffer
        const std::vector<int>& rr = cs->row_start[(size_t)r];
        const int q0 = rr[(size_t)ty0], n = rr[(size_t)ty1] - rr[(size_t)ty0];
        if (n <= 0) continue;
        const size_t at = (size_t)cs->offset[(size_t)r] + (size_t)q0;
        k_scatter_owned<<<(n * 3 + 255) / 256, 256, 0, st>>>(cs->stage.p + at * 3, cs->pix_others.p + at, n, fb);
      }
      CCUDA(cudaGetLastError());
    }
  }
  if (cs->rank == 0 && fb_out) {
    const size_t r0 = (size_t)std::min(ctx->yres, ty0 * 32) * ctx->xres * 3, r1 = (size_t)std::min(ctx->yres, ty1 * 32) * ctx->xres * 3;
    if (r1 > r0) CCUDA(cudaMemcpyAsync(fb_out + r0, fb + r0, (r1 - r0) * sizeof(float), cudaMemcpyDeviceToHost, st));  // a pageable destination is staged by the driver
  }
  if (sync) CCUDA(cudaStreamSynchronize(st));
  return VG_OK;
}
}  // namespace vg

static int comm_gather_rows_d2h_only(vg_ctx* ctx, cudaStream_t st, float* fb_out) {
  CommState* cs = ctx->comm;
  if (cs->rank != 0 || !fb_out) return VG_OK;
  float* fb = nullptr;
  int rc = render_fb_device(ctx, &fb);
  if (rc != VG_OK) return rc;
  CCUDA(cudaMemcpyAsync(fb_out, fb, (size_t)ctx->xres * ctx->yres * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
  return VG_OK;
}

extern "C" int vg_gather_frame(vg_ctx* ctx, float* fb_out) {
  if (!ctx) return VG_ERR_INVALID;
  std::lock_guard<std::mutex> lock(ctx->mu);
  CommState* cs = ctx->comm;
  if (!cs) return ctx->fail(VG_ERR_INVALID, "vg_gather_frame: no communicator (vg_comm_init)");
  CCUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  cudaEventRecord(cs->e0, st);
  int rc = comm_gather_rows(ctx, st, 0, (ctx->yres + 31) / 32, nullptr, false);
  if (rc != VG_OK) return rc;
  cudaEventRecord(cs->e1, st);
  rc = comm_gather_rows_d2h_only(ctx, st, fb_out);
  if (rc != VG_OK) return rc;
  CCUDA(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, cs->e0, cs->e1);
  cs->last_ms = ms;
  cs->gathers++;
  ctx->stats.kernel_launches += cs->world > 1 ? 1 : 0;
  ctx->stats.gather_ms = ms;
  return VG_OK;
}

extern "C" int vg_render_frame(vg_ctx* ctx, const uint64_t* table, int64_t npix, int iter_begin, int iter_end, int clear_first, float* fb_out) {
  if (!ctx) return VG_ERR_INVALID;
  std::lock_guard<std::mutex> lock(ctx->mu);
  CCUDA(cudaSetDevice(ctx->device));
  return render_frame(ctx, table, npix, iter_begin, iter_end, clear_first, fb_out);
}

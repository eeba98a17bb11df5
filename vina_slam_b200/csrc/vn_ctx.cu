// Context management and the low-level entry points of the C ABI
// (include/vina_b200.h): memory pools in HBM, per-stage kernel launches.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>
#include "vn_ctx.h"
#include <cstdlib>

#define CU(call)                                               \
  do                                                           \
  {                                                            \
    int _r = vn_check_cuda(ctx, (call), #call);                \
    if (_r) return _r;                                         \
  } while (0)

int vn_fail(vina_ctx* c, int code, const char* fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->err = buf;
  return code;
}

int vn_check_cuda(vina_ctx* c, cudaError_t e, const char* what)
{
  if (e == cudaSuccess) return VINA_OK;
  return vn_fail(c, VINA_E_CUDA, "CUDA error %s (%d) at %s", cudaGetErrorString(e), (int)e, what);
}

static int status_to_code(vina_ctx* ctx, int st)
{
  if (st == 0) return VINA_OK;
  if (st & VN_ST_UNSORTED) return vn_fail(ctx, VINA_E_ORDER, "scan is not sorted by time offset (status 0x%x)", st);
  if (st & VN_ST_SPIN) return vn_fail(ctx, VINA_E_CUDA, "device spin-wait limit hit (status 0x%x)", st);
  return vn_fail(ctx, VINA_E_CAPACITY,
                 "device pool exhausted or key out of range (status 0x%x: 1=hash 2=nodes 4=window arena 8=fixed pool "
                 "16=key range 128=downsample table)",
                 st);
}

int vn_check_status(vina_ctx* ctx)
{
  CU(cudaMemcpyAsync(ctx->h_status, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return status_to_code(ctx, *ctx->h_status);
}

extern "C" void vina_config_default(vina_config* c)
{
  memset(c, 0, sizeof(*c));
  c->voxel_size = 1.0;
  c->min_eigen_value = 0.0025;
  for (int i = 0; i < 4; i++) c->plane_eigen_value_thre[i] = 1.0;
  c->min_point[0] = 20;
  c->min_point[1] = 20;
  c->min_point[2] = 15;
  c->min_point[3] = 10;
  c->dept_err = 0.02;
  c->beam_err = 0.05;
  c->down_size = 0.1;
  c->ext_R[0] = c->ext_R[4] = c->ext_R[8] = 1.0;
  c->cov_gyr = 0.01;
  c->cov_acc = 1.0;
  c->rdw_gyr = 1e-4;
  c->rdw_acc = 1e-4;
  c->max_layer = 2;
  c->max_points = 100;
  c->win_size = 10;
  c->thread_num = 5;
}

template <typename T>
static cudaError_t dalloc(T** p, size_t count, bool zero = true)
{
  cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
  if (e != cudaSuccess) return e;
  if (zero) e = cudaMemset(*p, 0, count * sizeof(T));
  return e;
}

extern "C" int vina_ctx_create(const vina_config* cfg_in, vina_ctx** out)
{
  if (!cfg_in || !out) return VINA_E_ARG;
  *out = nullptr;
  vina_ctx* ctx = new vina_ctx();
  ctx->cfg = *cfg_in;
  vina_config& cfg = ctx->cfg;
  memset(&ctx->tm, 0, sizeof(ctx->tm));
  memset(&ctx->map, 0, sizeof(ctx->map));
  memset(&ctx->ins, 0, sizeof(ctx->ins));
  memset(&ctx->layers, 0, sizeof(ctx->layers));
  memset(&ctx->peers, 0, sizeof(ctx->peers));
  memset(ctx->pv, 0, sizeof(ctx->pv));
  if (cfg.win_size < 1 || cfg.win_size > VINA_MAX_WIN || cfg.max_layer < 0 || cfg.max_layer > 3 ||
      cfg.voxel_size <= 0 || cfg.thread_num < 1)
  {
    delete ctx;
    return VINA_E_ARG;
  }
  if (cfg.max_scan_points <= 0) cfg.max_scan_points = 300000;
  if (cfg.max_nodes <= 0) cfg.max_nodes = 1 << 20;
  if (cfg.hash_capacity_log2 <= 0) cfg.hash_capacity_log2 = 21;
  if (cfg.fix_pool_points <= 0) cfg.fix_pool_points = 16ll << 20;
  if (cfg.win_pool_points <= 0) cfg.win_pool_points = 4ll * cfg.max_scan_points;
  if (cfg.max_points <= 0) cfg.max_points = 100;
  // bounds before anything is allocated: 1u << 32 is undefined, the down-sampling table is sized 2 x points in 32 bits
  if (cfg.hash_capacity_log2 < 10 || cfg.hash_capacity_log2 > 31 || cfg.max_scan_points > (1 << 29) ||
      cfg.max_nodes > (1 << 30) || cfg.device < 0)
  {
    delete ctx;
    return VINA_E_ARG;
  }

  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
  {
    // no CPU fallback: the product path needs a CUDA device
    fprintf(stderr, "vina_b200: no CUDA device (%s); there is no CPU fallback\n", cudaGetErrorString(e));
    delete ctx;
    return VINA_E_CUDA;
  }
  if (cfg.device >= ndev)
  {
    delete ctx;
    return VINA_E_ARG;
  }
  ctx->device = cfg.device;
  *out = ctx;  // from here on errors keep the ctx so that vina_last_error works
  CU(cudaSetDevice(ctx->device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, ctx->device));
  ctx->sm_count = prop.multiProcessorCount;
  CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  ctx->own_stream = true;
  for (int i = 0; i < 16; i++) CU(cudaEventCreate(&ctx->ev[i]));

  const int cap = cfg.max_scan_points;
  ctx->cap_points = cap;
  CU(dalloc(&ctx->d_scan, cap));
  CU(dalloc(&ctx->d_down, cap));
  CU(dalloc(&ctx->d_n_down, 1));
  CU(cudaHostAlloc((void**)&ctx->h_n_down, sizeof(int), cudaHostAllocDefault));
  for (int w = 0; w < 2; w++)
  {
    // one allocation per pointVar set: 9 contiguous arrays (k_iekf addresses them as base + k * cap)
    double* base = nullptr;
    CU(dalloc(&base, (size_t)9 * cap));
    for (int k = 0; k < 3; k++) ctx->pv[w].p[k] = base + (size_t)k * cap;
    for (int k = 0; k < 6; k++) ctx->pv[w].v[k] = base + (size_t)(3 + k) * cap;
  }
  CU(dalloc(&ctx->d_cache, cap));
  CU(dalloc(&ctx->d_poses, 1));
  CU(cudaHostAlloc((void**)&ctx->h_poses, sizeof(DeskewPoses), cudaHostAllocDefault));
  // down-sampling table: >= 2x points
  unsigned int dslots = 1;
  while (dslots < 2u * (unsigned)cap) dslots <<= 1;
  ctx->dmask = dslots - 1;
  CU(dalloc(&ctx->d_dtab, dslots, false));
  CU(dalloc(&ctx->d_slot_of, cap));
  CU(dalloc(&ctx->d_flag, cap));
  CU(dalloc(&ctx->d_scanbuf, cap));
  CU(dalloc(&ctx->d_block_sums, 1024));
  launch_down_init(ctx->stream, ctx->d_dtab, dslots);
  // IEKF
  CU(dalloc(&ctx->d_partials, ((size_t)cap / 128 + 64) * VN_IEKF_NACC));
  CU(dalloc(&ctx->d_ticket, 1));
  CU(dalloc(&ctx->d_emit_counts, ctx->sm_count));
  CU(dalloc(&ctx->d_emit_bar, 2));
  CU(cudaHostAlloc((void**)&ctx->h_down_pub, 64, cudaHostAllocMapped));
  CU(cudaHostGetDevicePointer((void**)&ctx->d_down_pub, ctx->h_down_pub, 0));
  memset(ctx->h_down_pub, 0, 64);
  if (const char* e = getenv("VINA_FRONT_FUSED")) ctx->front_fused = atoi(e) != 0;
  CU(dalloc(&ctx->d_loop_bar, 2));
  CU(dalloc(&ctx->d_loop_partials, (size_t)2 * VN_IEKF_NACC * ctx->sm_count));
  if (const char* e = getenv("VINA_IEKF_LOOP")) ctx->iekf_loop = atoi(e) != 0;
  CU(cudaHostAlloc((void**)&ctx->h_result, 64 * sizeof(double), cudaHostAllocMapped));
  CU(cudaHostGetDevicePointer((void**)&ctx->d_result, ctx->h_result, 0));
  memset(ctx->h_result, 0, 64 * sizeof(double));
  CU(dalloc(&ctx->d_iekf, 1));
  CU(cudaHostAlloc((void**)&ctx->h_iekf, sizeof(IekfDev), cudaHostAllocDefault));
  memset(ctx->h_iekf, 0, sizeof(IekfDev));
  CU(cudaHostAlloc((void**)&ctx->h_pub, sizeof(IekfDev), cudaHostAllocMapped));
  CU(cudaHostGetDevicePointer((void**)&ctx->d_pub, ctx->h_pub, 0));
  CU(cudaHostAlloc((void**)&ctx->h_pub_flag, 64, cudaHostAllocMapped));
  CU(cudaHostGetDevicePointer((void**)&ctx->d_pub_flag, ctx->h_pub_flag, 0));
  *ctx->h_pub_flag = 0ull;
  CU(cudaEventCreateWithFlags(&ctx->ev_poses, cudaEventDisableTiming));
  CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&ctx->ev_scan_up, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&ctx->ev_scan_rd, cudaEventDisableTiming));
  for (int i = 0; i < 4; i++) CU(cudaEventCreateWithFlags(&ctx->ev_chunk[i], cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&ctx->ev_step_begin, cudaEventDisableTiming));
  CU(cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
  ctx->trace = getenv("VINA_TRACE") != nullptr;
  if (ctx->trace)
    for (int i = 0; i < 8; i++) CU(cudaEventCreate(&ctx->tr_ev[i]));
  CU(cudaEventCreateWithFlags(&ctx->ev_collect_fork, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&ctx->ev_collect_done, cudaEventDisableTiming));
  if (const char* e = getenv("VINA_EARLY_COLLECT")) ctx->early_collect = atoi(e) != 0;
  CU(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
  CU(dalloc(&ctx->d_status, 1));
  CU(cudaHostAlloc((void**)&ctx->h_status, sizeof(int), cudaHostAllocDefault));

  // map
  MapView& M = ctx->map;
  ctx->hash_slots = 1u << cfg.hash_capacity_log2;
  M.hmask = ctx->hash_slots - 1;
  CU(dalloc(&M.slots, ctx->hash_slots, false));
  M.max_nodes = cfg.max_nodes;
  CU(dalloc(&M.hot, (size_t)cfg.max_nodes));
  CU(dalloc(&M.cold, (size_t)cfg.max_nodes));
  CU(dalloc(&M.node_count, 1));
  CU(dalloc(&M.root_count, 1));
  M.win_cap = cfg.win_pool_points;
  for (int s = 0; s < cfg.win_size; s++) CU(dalloc(&M.win_pool[s], (size_t)cfg.win_pool_points, false));
  CU(dalloc(&M.win_cursor, VINA_MAX_WIN));
  M.fix_cap = cfg.fix_pool_points;
  CU(dalloc(&M.fix_pool, (size_t)cfg.fix_pool_points, false));
  // chain blocks of 15 segments (128 B each): one per leaf that keeps fixed points, plus one per 15 folds
  M.fixseg_cap = (int)(cfg.fix_pool_points / 16 > (1 << 28) ? (1 << 28) : cfg.fix_pool_points / 16);
  if (M.fixseg_cap < 4096) M.fixseg_cap = 4096;
  CU(dalloc(&M.fix_segs, (size_t)M.fixseg_cap, false));
  CU(dalloc(&M.fix_cursor, 1));
  CU(dalloc(&M.fixseg_cursor, 1));
  // free stacks of the map pruning (4 B per node / chain block)
  CU(dalloc(&M.free_nodes, (size_t)cfg.max_nodes, false));
  CU(dalloc(&M.free_count, 1));
  CU(dalloc(&M.free_segs, (size_t)M.fixseg_cap, false));
  CU(dalloc(&M.free_seg_count, 1));
  CU(dalloc(&ctx->d_prune, 4));
  M.jour = 0.0;
  CU(dalloc(&M.slide_list[0], (size_t)cfg.max_nodes));
  CU(dalloc(&M.slide_list[1], (size_t)cfg.max_nodes));
  CU(dalloc(&M.slide_count, 2));
  M.slide_cur = 0;
  M.status = ctx->d_status;
  M.voxel_size = cfg.voxel_size;
  M.min_eigen_value = cfg.min_eigen_value;
  for (int i = 0; i < 4; i++)
  {
    M.thre[i] = 1.0 / cfg.plane_eigen_value_thre[i];  // node.cpp:256-259
    M.min_point[i] = cfg.min_point[i];
  }
  M.max_layer = cfg.max_layer;
  M.max_points = cfg.max_points;
  M.win_size = cfg.win_size;
  M.thread_num = cfg.thread_num;
  for (int i = 0; i < VINA_MAX_WIN; i++) M.mp[i] = i;  // node.cpp:431-435
  launch_map_init(ctx->stream, M, ctx->hash_slots);

  InsertScratch& S = ctx->ins;
  for (int k = 0; k < 3; k++) CU(dalloc(&S.pw[k], cap));
  for (int k = 0; k < 6; k++) CU(dalloc(&S.vw[k], cap));
  CU(dalloc(&S.root_of, cap));
  CU(dalloc(&S.leaf_of, cap));
  CU(dalloc(&S.rank_of, cap));
  CU(dalloc(&S.touched, cap));
  CU(dalloc(&S.counters, 8));
  S.counters_alt = S.counters + 4;
  CU(dalloc(&S.idx, cap));
  S.stamp = 0;
  for (int l = 0; l < 4; l++) CU(dalloc(&ctx->layers.list[l], (size_t)cfg.max_nodes));
  CU(dalloc(&ctx->layers.split, (size_t)cfg.max_nodes));
  CU(cudaMemset(ctx->layers.split, 0xFF, (size_t)cfg.max_nodes * sizeof(int)));  // k_split's queue: -1 = slot not written yet
  CU(dalloc(&ctx->layers.count, 16));
  ctx->layers.count_alt = ctx->layers.count + 8;
  CU(dalloc(&ctx->layers.snap, 4));
  if (const char* e = getenv("VINA_SPLIT_OVERLAP")) ctx->split_overlap = atoi(e) != 0;
  CU(cudaStreamSynchronize(ctx->stream));
  CU(cudaDeviceSynchronize());  // the zero-fills of dalloc ran on the legacy stream (see ensure_debug)
  CU(cudaGetLastError());
  return VINA_OK;
}

#ifdef VINA_SPLIT_TRACE
void vn_split_trace_dump();  // map_kernels.cu
#endif
extern "C" void vina_ctx_destroy(vina_ctx* ctx)
{
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->odom) odom_host_destroy(ctx->odom);
  cudaFree(ctx->d_scan);
  cudaFree(ctx->d_down);
  cudaFree(ctx->d_n_down);
  cudaFreeHost(ctx->h_n_down);
  for (int w = 0; w < 2; w++) cudaFree(ctx->pv[w].p[0]);
  cudaFree(ctx->d_cache);
  cudaFree(ctx->d_poses);
  cudaFreeHost(ctx->h_poses);
  cudaFree(ctx->d_dtab);
  cudaFree(ctx->d_slot_of);
  cudaFree(ctx->d_flag);
  cudaFree(ctx->d_scanbuf);
  cudaFree(ctx->d_block_sums);
  cudaFree(ctx->d_partials);
  cudaFree(ctx->d_ticket);
  cudaFree(ctx->d_emit_counts);
  cudaFree(ctx->d_emit_bar);
  if (ctx->h_down_pub) cudaFreeHost(ctx->h_down_pub);
  cudaFree(ctx->d_loop_bar);
  cudaFree(ctx->d_loop_partials);
  cudaFreeHost(ctx->h_result);
  cudaFree(ctx->d_iekf);
  cudaFreeHost(ctx->h_iekf);
  cudaFreeHost(ctx->h_pub);
  cudaFreeHost(ctx->h_pub_flag);
  if (ctx->ev_poses) cudaEventDestroy(ctx->ev_poses);
  if (ctx->ev_scan_up) cudaEventDestroy(ctx->ev_scan_up);
  if (ctx->ev_scan_rd) cudaEventDestroy(ctx->ev_scan_rd);
  if (ctx->copy_stream)
  {
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
  }
  if (ctx->side_stream)
  {
    cudaStreamSynchronize(ctx->side_stream);
    cudaStreamDestroy(ctx->side_stream);
  }
#ifdef VINA_SPLIT_TRACE
  if (ctx->trace) vn_split_trace_dump();
#endif
  if (ctx->trace && ctx->tr_n > 0)
  {
    fprintf(stderr, "[vina trace] %d overlapped steps; host us (propagate, front+iekf enqueued, down count, map enqueued, "
                    "iterate arrived):", ctx->tr_n);
    for (int i = 1; i <= 5; i++) fprintf(stderr, " %.1f", ctx->tr_host_us[i] / ctx->tr_n);
    fprintf(stderr, "; device us since deskew start (iekf done, side stream joined, insert, recut, margi):");
    for (int i = 1; i <= 5; i++) fprintf(stderr, " %.1f", ctx->tr_dev_us[i] / ctx->tr_n);
    fprintf(stderr, "\n");
  }
  for (int i = 0; i < 8; i++)
    if (ctx->tr_ev[i]) cudaEventDestroy(ctx->tr_ev[i]);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  cudaFree(ctx->d_ba);
  cudaFree(ctx->d_ba_n);
  cudaFree(ctx->d_ba_partial);
  cudaFree(ctx->d_ba_out);
  cudaFree(ctx->d_ba_lam);
  cudaFreeHost(ctx->h_ba_out);
  cudaFreeHost(ctx->h_ba_flag);
  cudaFree(ctx->d_ba_ticket);
  cudaFree(ctx->d_status);
  cudaFreeHost(ctx->h_status);
  for (void* m : ctx->p2p_opened)
    if (m) cudaIpcCloseMemHandle(m);
  cudaFree(ctx->p2p_inbox);
  cudaFree(ctx->p2p_ctrl);
  cudaFree(ctx->d_n_recv);
  cudaFree(ctx->d_p2p_base);
  cudaFree(ctx->d_sh_owner);
  cudaFree(ctx->d_sh_hist);
  cudaFree(ctx->d_sh_counts);
  cudaFreeHost(ctx->h_sh_counts);
  cudaFree(ctx->dbg.keys);
  cudaFree(ctx->dbg.codes);
  cudaFree(ctx->dbg.flags);
  cudaFree(ctx->dbg.sigma);
  MapView& M = ctx->map;
  cudaFree(M.slots);
  cudaFree(M.hot);
  cudaFree(M.cold);
  cudaFree(M.node_count);
  cudaFree(M.root_count);
  for (int s = 0; s < VINA_MAX_WIN; s++) cudaFree(M.win_pool[s]);
  cudaFree(M.win_cursor);
  cudaFree(M.fix_pool);
  cudaFree(M.fix_segs);
  cudaFree(M.fix_cursor);
  cudaFree(M.fixseg_cursor);
  cudaFree(ctx->d_raw);
  for (int k = 0; k < 2; k++)
  {
    cudaFree(ctx->d_fkey[k]);
    cudaFree(ctx->d_fidx[k]);
  }
  cudaFree(ctx->d_fhist);
  cudaFree(ctx->d_fbkt);
  if (ctx->h_front_pub) cudaFreeHost(ctx->h_front_pub);
  cudaFree(ctx->d_fpairs);
  cudaFree(ctx->d_fcnt);
  if (ctx->h_fcnt) cudaFreeHost(ctx->h_fcnt);
  cudaFree(M.free_nodes);
  cudaFree(M.free_count);
  cudaFree(M.free_segs);
  cudaFree(M.free_seg_count);
  cudaFree(ctx->d_prune);
  cudaFree(M.slide_list[0]);
  cudaFree(M.slide_list[1]);
  cudaFree(M.slide_count);
  InsertScratch& S = ctx->ins;
  for (int k = 0; k < 3; k++) cudaFree(S.pw[k]);
  for (int k = 0; k < 6; k++) cudaFree(S.vw[k]);
  cudaFree(S.root_of);
  cudaFree(S.leaf_of);
  cudaFree(S.rank_of);
  cudaFree(S.touched);
  cudaFree(S.counters < S.counters_alt ? S.counters : S.counters_alt);
  cudaFree(S.idx);
  for (int k = 0; k < 2; k++) cudaFree(ctx->d_tree[k]);
  cudaFree(ctx->d_init_ds);
  cudaFree(ctx->d_init_dir);
  cudaFree(ctx->d_init_part);
  if (ctx->h_init_sums) cudaFreeHost(ctx->h_init_sums);
  for (int l = 0; l < 4; l++) cudaFree(ctx->layers.list[l]);
  cudaFree(ctx->layers.split);
  cudaFree(ctx->layers.snap);
  cudaFree(ctx->layers.count < ctx->layers.count_alt ? ctx->layers.count : ctx->layers.count_alt);
  for (int i = 0; i < 16; i++)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (cudaEvent_t e : ctx->iekf_ev) cudaEventDestroy(e);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* vina_last_error(vina_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

extern "C" int vina_ctx_set_stream(vina_ctx* ctx, void* cuda_stream)
{
  if (!ctx) return VINA_E_ARG;
  CU(cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)cuda_stream;
  ctx->own_stream = false;
  // uploads run on the copy stream: the new compute stream has to see the last one as well
  CU(cudaStreamSynchronize(ctx->copy_stream));
  return VINA_OK;
}

extern "C" int vina_ctx_sync(vina_ctx* ctx)
{
  if (!ctx) return VINA_E_ARG;
  return vn_check_status(ctx);
}

// ---------------------------------------------------------------------------
// every entry that reads d_scan marks the stream position after its last reader, so that the next upload
// (on the copy stream) only waits for those kernels and not for the map update enqueued behind them
static int mark_scan_read(vina_ctx* ctx)
{
  CU(cudaEventRecord(ctx->ev_scan_rd, ctx->stream));
  ctx->scan_rd_valid = true;
  return VINA_OK;
}

int vn_mark_scan_read(vina_ctx* ctx) { return mark_scan_read(ctx); }

extern "C" int vina_set_overlap(vina_ctx* ctx, int on)
{
  if (!ctx) return VINA_E_ARG;
  ctx->overlap = on != 0;
  return VINA_OK;
}

extern "C" int vina_set_iekf_loop(vina_ctx* ctx, int on)
{
  if (!ctx) return VINA_E_ARG;
  ctx->iekf_loop = on != 0;
  return VINA_OK;
}

extern "C" int vina_scan_upload(vina_ctx* ctx, const float* xyzt, int n)
{
  if (!ctx || !xyzt || n < 0) return VINA_E_ARG;
  if (n > ctx->cap_points) return vn_fail(ctx, VINA_E_CAPACITY, "scan of %d points > max_scan_points %d", n, ctx->cap_points);
  // the copy runs on its own stream: it starts as soon as the previous readers of d_scan are done and overlaps
  // whatever else is still queued on the compute stream (the previous scan's map update)
  if (ctx->scan_rd_valid) CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_scan_rd, 0));
  CU(cudaMemcpyAsync(ctx->d_scan, xyzt, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, ctx->copy_stream));
  CU(cudaEventRecord(ctx->ev_scan_up, ctx->copy_stream));
  CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_scan_up, 0));
  ctx->n_scan = n;
  ctx->front_valid = false;
  return VINA_OK;
}

extern "C" int vina_set_upload_ordered(vina_ctx* ctx, int on)
{
  if (!ctx) return VINA_E_ARG;
  ctx->upload_ordered = on != 0;
  return VINA_OK;
}

// the upload of vina_odom_step: up to four chunks on the copy stream, an event behind each; the compute stream does
// NOT wait here - vn_deskew_var_init follows the chunks
int vn_scan_upload_chunked(vina_ctx* ctx, const float* xyzt, int n)
{
  if (n > ctx->cap_points) return vn_fail(ctx, VINA_E_CAPACITY, "scan of %d points > max_scan_points %d", n, ctx->cap_points);
  if (ctx->scan_rd_valid) CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_scan_rd, 0));
  if (ctx->upload_ordered)
  {
    CU(cudaEventRecord(ctx->ev_step_begin, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_step_begin, 0));
  }
  // two chunks, 3/4 + 1/4: every extra copy costs ~5 us of copy time, so the aim is only to leave a short tail of the
  // deskew behind the last byte (one 3.84 MB copy: 76 us at 50 GB/s; its deskew: 24 us)
  const int chunks = n >= 32768 ? 2 : 1;
  int first = 0;
  for (int c = 0; c < chunks; c++)
  {
    int last = c + 1 == chunks ? n : ((int)(((long long)n * 3) / 4) & ~255);
    CU(cudaMemcpyAsync(ctx->d_scan + first, xyzt + 4 * (size_t)first, (size_t)(last - first) * sizeof(float4),
                       cudaMemcpyHostToDevice, ctx->copy_stream));
    CU(cudaEventRecord(ctx->ev_chunk[c], ctx->copy_stream));
    ctx->chunk_end[c] = last;
    first = last;
  }
  ctx->upload_chunks = chunks;
  ctx->n_scan = n;
  ctx->front_valid = false;
  return VINA_OK;
}

// whoever consumes d_scan other than the chunk-following deskew first lets the compute stream wait for the whole upload
static int settle_upload(vina_ctx* ctx)
{
  if (ctx->upload_chunks > 0)
  {
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[ctx->upload_chunks - 1], 0));
    ctx->upload_chunks = 0;
  }
  return VINA_OK;
}
int vn_settle_upload(vina_ctx* ctx) { return settle_upload(ctx); }

extern "C" int vina_scan_upload_device(vina_ctx* ctx, const void* d_xyzt, int n)
{
  if (!ctx || !d_xyzt || n < 0) return VINA_E_ARG;
  if (n > ctx->cap_points) return vn_fail(ctx, VINA_E_CAPACITY, "scan of %d points > max_scan_points %d", n, ctx->cap_points);
  CU(cudaMemcpyAsync(ctx->d_scan, d_xyzt, (size_t)n * sizeof(float4), cudaMemcpyDeviceToDevice, ctx->stream));
  // a later vina_scan_upload (copy stream) must order behind this write of d_scan
  CU(cudaEventRecord(ctx->ev_scan_rd, ctx->stream));
  ctx->scan_rd_valid = true;
  ctx->n_scan = n;
  ctx->front_valid = false;  // the buffer no longer holds the scan vina_scan_prepare left there
  return VINA_OK;
}

// ---------------------------------------------------------------------------
// scan front end: decoder keep rule + pcl_handler (lidar_pointcloud_decoder.cpp:70; lidar_decoder.cpp:16-34)
static int ensure_front(vina_ctx* ctx)
{
  if (ctx->d_fcnt) return VINA_OK;
  const size_t cap = ctx->cap_points;
  CU(dalloc(&ctx->d_raw, cap, false));
  for (int k = 0; k < 2; k++)
  {
    CU(dalloc(&ctx->d_fkey[k], cap, false));
    CU(dalloc(&ctx->d_fidx[k], cap, false));
  }
  CU(dalloc(&ctx->d_fhist, 256 * (cap / 2048 + 1)));
  CU(dalloc(&ctx->d_fbkt, 3 * 1024 + 8));
  CU(cudaHostAlloc((void**)&ctx->h_front_pub, 64, cudaHostAllocMapped));
  CU(cudaHostGetDevicePointer((void**)&ctx->d_front_pub, ctx->h_front_pub, 0));
  memset(ctx->h_front_pub, 0, 64);
  CU(dalloc(&ctx->d_fpairs, cap, false));
  CU(dalloc(&ctx->d_fcnt, 4));
  CU(cudaHostAlloc((void**)&ctx->h_fcnt, 4 * sizeof(int), cudaHostAllocDefault));
  CU(cudaDeviceSynchronize());
  return VINA_OK;
}

static int front_run(vina_ctx* ctx, const float4* d_raw, int n, int point_filter_num, double blind2, int* n_out,
                     float* t_last)
{
  ctx->front_valid = false;
  int kept = 0, keep = 0;
  float tl = 0.f;
  if (n > 0)
  {
    // one bucket pass by time + a shared-memory sort per bucket; the radix passes only if a bucket overflowed
    // (VINA_FRONT_RADIX=1 forces them)
    static const bool force_radix = getenv("VINA_FRONT_RADIX") && atoi(getenv("VINA_FRONT_RADIX")) != 0;
    if (!force_radix)
    {
      const unsigned long long seq = ++ctx->front_seq;
      ctx->launches += launch_front_prepare_buckets(ctx->stream, d_raw, n, point_filter_num, blind2, ctx->d_fidx[0], ctx->d_fbkt,
                                                    ctx->d_fpairs, ctx->d_scan, ctx->d_front_pub, seq);
      // the counters arrive through mapped memory (written by the block that finishes last)
      volatile unsigned long long* pub = ctx->h_front_pub;
      for (long spins = 0; pub[0] != seq; spins++)
        if ((spins & 0xfff) == 0xfff)
        {
          cudaError_t e = cudaStreamQuery(ctx->stream);
          if (e == cudaSuccess)
          {
            if (pub[0] == seq) break;
            return vn_fail(ctx, VINA_E_CUDA, "the scan front end finished without publishing its counters");
          }
          if (e != cudaErrorNotReady) return vn_check_cuda(ctx, e, "scan front end");
        }
      __sync_synchronize();
      for (int k = 0; k < 4; k++) ctx->h_fcnt[k] = (int)(unsigned int)pub[1 + k];
    }
    if (force_radix || ctx->h_fcnt[3])
    {
      ctx->launches += launch_front_prepare(ctx->stream, d_raw, n, point_filter_num, blind2, ctx->d_fkey, ctx->d_fidx,
                                            ctx->d_fhist, ctx->d_fcnt, ctx->d_scan, reinterpret_cast<float*>(ctx->d_fcnt + 2));
      CU(cudaMemcpyAsync(ctx->h_fcnt, ctx->d_fcnt, 3 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
    }
    kept = ctx->h_fcnt[0];
    keep = ctx->h_fcnt[1];
    memcpy(&tl, &ctx->h_fcnt[2], 4);
  }
  if (kept == 0)
  {
    // pcl_handler's stand-in for an empty cloud (lidar_decoder.cpp:16-27): two points at the origin, 0 s and 0.09 s
    const float dummy[8] = { 0, 0, 0, 0, 0, 0, 0, 0.09f };
    if (ctx->cap_points < 2) return vn_fail(ctx, VINA_E_CAPACITY, "max_scan_points < 2");
    CU(cudaMemcpyAsync(ctx->d_scan, dummy, sizeof(dummy), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    keep = 2;
    tl = 0.09f;
  }
  else if (keep == 0)
    return vn_fail(ctx, VINA_E_ARG, "no point of the scan lies within 0.11 s (the reference pops an empty cloud here)");
  ctx->n_scan = keep;
  ctx->front_t_last = tl;
  ctx->front_valid = true;
  if (n_out) *n_out = keep;
  if (t_last) *t_last = tl;
  return VINA_OK;
}

extern "C" int vina_scan_prepare(vina_ctx* ctx, const float* xyzt, int n, int point_filter_num, double blind2, int* n_out,
                                 float* t_last)
{
  if (!ctx || (!xyzt && n > 0) || n < 0 || point_filter_num < 1) return VINA_E_ARG;
  if (n > ctx->cap_points) return vn_fail(ctx, VINA_E_CAPACITY, "scan of %d points > max_scan_points %d", n, ctx->cap_points);
  int r = ensure_front(ctx);
  if (r) return r;
  if (n > 0) CU(cudaMemcpyAsync(ctx->d_raw, xyzt, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
  return front_run(ctx, ctx->d_raw, n, point_filter_num, blind2, n_out, t_last);
}

extern "C" int vina_scan_prepare_device(vina_ctx* ctx, const void* d_xyzt, int n, int point_filter_num, double blind2,
                                        int* n_out, float* t_last)
{
  if (!ctx || (!d_xyzt && n > 0) || n < 0 || point_filter_num < 1) return VINA_E_ARG;
  if (n > ctx->cap_points) return vn_fail(ctx, VINA_E_CAPACITY, "scan of %d points > max_scan_points %d", n, ctx->cap_points);
  int r = ensure_front(ctx);
  if (r) return r;
  return front_run(ctx, reinterpret_cast<const float4*>(d_xyzt), n, point_filter_num, blind2, n_out, t_last);
}

extern "C" int vina_deskew(vina_ctx* ctx, const vina_imu_pose* poses, int m, const double R_end[9],
                           const double p_end[3])
{
  if (!ctx || !poses || m < 0 || !R_end || !p_end) return VINA_E_ARG;
  if (m > VINA_MAX_POSES) return vn_fail(ctx, VINA_E_CAPACITY, "%d IMU poses > VINA_MAX_POSES", m);
  // the pinned staging buffer may still be in flight from the previous scan (never in steady state)
  if (ctx->poses_in_flight) CU(cudaEventSynchronize(ctx->ev_poses));
  DeskewPoses* P = ctx->h_poses;
  P->m = m;
  memcpy(P->pose, poses, (size_t)m * sizeof(vina_imu_pose));
  memcpy(P->R_end, R_end, 72);
  memcpy(P->p_end, p_end, 24);
  memcpy(P->ext_R, ctx->cfg.ext_R, 72);
  memcpy(P->ext_t, ctx->cfg.ext_t, 24);
  CU(cudaMemcpyAsync(ctx->d_poses, P, sizeof(DeskewPoses), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaEventRecord(ctx->ev_poses, ctx->stream));
  ctx->poses_in_flight = true;
  launch_deskew(ctx->stream, ctx->d_scan, ctx->n_scan, ctx->d_poses, ctx->d_status);
  ctx->launches += 1;
  return mark_scan_read(ctx);
}

extern "C" int vina_scan_download(vina_ctx* ctx, float* xyzt, int cap)
{
  if (!ctx || !xyzt) return VINA_E_ARG;
  if (cap < ctx->n_scan) return VINA_E_ARG;
  CU(cudaMemcpyAsync(xyzt, ctx->d_scan, (size_t)ctx->n_scan * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
  int r = mark_scan_read(ctx);
  if (r) return r;
  r = vn_check_status(ctx);
  if (r) return r;
  return ctx->n_scan;
}

static VarInitParams var_init_params(const vina_ctx* ctx);
static int run_downsample(vina_ctx* ctx, double size)
{
  // (the per-scan step lets the emission do the var_init of the emitted set as well: ctx->down_fuse_var_init)
  const VarInitParams prm = var_init_params(ctx);
  int k = launch_downsample(ctx->stream, ctx->d_scan, ctx->n_scan, size, ctx->d_dtab, ctx->dmask, ctx->d_slot_of,
                            ctx->d_flag, ctx->d_scanbuf, ctx->d_block_sums, ctx->d_n_down, ctx->d_down, ctx->d_status,
                            ctx->d_down_pub, ++ctx->down_seq, ctx->down_fuse_var_init ? &ctx->pv[1] : nullptr,
                            ctx->down_fuse_var_init ? &prm : nullptr);
  if (k < 0) return vn_fail(ctx, VINA_E_CAPACITY, "scan too large for the down-sampling scan kernels");
  ctx->launches += k;
  ctx->n_down_pending = true;
  ctx->n_down_mapped = k > 0;  // the count arrives through mapped memory (k_scan_sums), resolve_n_down polls for it
  if (k == 0)
  {
    ctx->n_down = 0;
    ctx->n_down_pending = false;
  }
  return mark_scan_read(ctx);
}

// resolve the pending down-sampled count (one stream sync) and apply the
// "< 2000 points -> down_size / 2" retry of local_mapping.cpp:399-403
static int resolve_n_down(vina_ctx* ctx)
{
  if (!ctx->n_down_pending) return VINA_OK;
  if (ctx->n_down_mapped)
  {
    // published by k_down_emit_all: poll the sequence number (no stream synchronisation; the IEKF loop behind it
    // keeps running)
    volatile unsigned long long* pub = ctx->h_down_pub;
    for (long spins = 0; pub[0] != ctx->down_seq; spins++)
      if ((spins & 0xfff) == 0xfff)
      {
        cudaError_t e = cudaStreamQuery(ctx->stream);
        if (e == cudaSuccess)
        {
          if (pub[0] == ctx->down_seq) break;
          return vn_fail(ctx, VINA_E_CUDA, "the down-sampling finished without publishing its count");
        }
        if (e != cudaErrorNotReady) return vn_check_cuda(ctx, e, "down-sampling");
      }
    __sync_synchronize();
    ctx->n_down = (int)pub[1];
    ctx->n_down_mapped = false;
    ctx->n_down_pending = false;
    return VINA_OK;
  }
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->n_down = *ctx->h_n_down;
  ctx->n_down_pending = false;
  return VINA_OK;
}

extern "C" int vina_downsample(vina_ctx* ctx)
{
  if (!ctx) return VINA_E_ARG;
  if (ctx->cfg.down_size < 0.001)
  {
    CU(cudaMemcpyAsync(ctx->d_down, ctx->d_scan, (size_t)ctx->n_scan * sizeof(float4), cudaMemcpyDeviceToDevice,
                       ctx->stream));
    ctx->n_down = ctx->n_scan;
    ctx->n_down_pending = false;
    return mark_scan_read(ctx);
  }
  return run_downsample(ctx, ctx->cfg.down_size);
}

// called by the pipeline once the count is needed on the host
int vn_finish_downsample(vina_ctx* ctx)
{
  int r = resolve_n_down(ctx);
  if (r) return r;
  ctx->down_retried = false;
  if (ctx->n_down < 2000 && ctx->cfg.down_size >= 0.001 && ctx->n_scan > 0)
  {
    ctx->down_retried = true;
    // retry with half the voxel size (local_mapping.cpp:399-403)
    r = run_downsample(ctx, ctx->cfg.down_size / 2);
    if (r) return r;
    r = resolve_n_down(ctx);
    if (r) return r;
  }
  return VINA_OK;
}

extern "C" int vina_down_count(vina_ctx* ctx)
{
  if (!ctx) return VINA_E_ARG;
  int r = vn_finish_downsample(ctx);
  if (r) return r;
  return ctx->n_down;
}

extern "C" int vina_down_upload(vina_ctx* ctx, const float* xyzt, int n)
{
  if (!ctx || !xyzt || n < 0) return VINA_E_ARG;
  if (n > ctx->cap_points) return vn_fail(ctx, VINA_E_CAPACITY, "cloud of %d points > max_scan_points", n);
  CU(cudaMemcpyAsync(ctx->d_down, xyzt, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
  ctx->n_down = n;
  ctx->n_down_pending = false;
  return VINA_OK;
}

extern "C" int vina_down_download(vina_ctx* ctx, float* xyzt, int cap)
{
  if (!ctx || !xyzt) return VINA_E_ARG;
  int r = resolve_n_down(ctx);
  if (r) return r;
  if (cap < ctx->n_down) return VINA_E_ARG;
  CU(cudaMemcpyAsync(xyzt, ctx->d_down, (size_t)ctx->n_down * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
  r = vn_check_status(ctx);
  if (r) return r;
  return ctx->n_down;
}

static VarInitParams var_init_params(const vina_ctx* ctx)
{
  VarInitParams prm;
  const float range_inc = (float)ctx->cfg.dept_err, degree_inc = (float)ctx->cfg.beam_err;  // point_utils.cpp:3
  prm.range_var = range_inc * range_inc;
  double s = sin((degree_inc)*M_PI / 180.0);
  prm.dir_var = s * s;
  memcpy(prm.ext_R, ctx->cfg.ext_R, 72);
  memcpy(prm.ext_t, ctx->cfg.ext_t, 24);
  return prm;
}

// vina_deskew + vina_var_init(ctx, 0) + the IEKF's leaf-cache reset as one kernel (the per-scan step)
int vn_deskew_var_init(vina_ctx* ctx, const vina_imu_pose* poses, int m, const double R_end[9], const double p_end[3])
{
  if (m > VINA_MAX_POSES) return vn_fail(ctx, VINA_E_CAPACITY, "%d IMU poses > VINA_MAX_POSES", m);
  if (ctx->poses_in_flight) CU(cudaEventSynchronize(ctx->ev_poses));
  DeskewPoses* P = ctx->h_poses;
  P->m = m;
  memcpy(P->pose, poses, (size_t)m * sizeof(vina_imu_pose));
  memcpy(P->R_end, R_end, 72);
  memcpy(P->p_end, p_end, 24);
  memcpy(P->ext_R, ctx->cfg.ext_R, 72);
  memcpy(P->ext_t, ctx->cfg.ext_t, 24);
  CU(cudaMemcpyAsync(ctx->d_poses, P, sizeof(DeskewPoses), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaEventRecord(ctx->ev_poses, ctx->stream));
  ctx->poses_in_flight = true;
  if (ctx->upload_chunks > 0)
  {
    // the scan is still arriving: one launch per chunk, each behind its chunk's copy
    int first = 0;
    for (int c = 0; c < ctx->upload_chunks; c++)
    {
      CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[c], 0));
      launch_deskew_var_init(ctx->stream, ctx->d_scan, ctx->n_scan, ctx->d_poses, ctx->d_status, ctx->pv[0], var_init_params(ctx),
                             ctx->d_cache, first, ctx->chunk_end[c]);
      first = ctx->chunk_end[c];
    }
    ctx->launches += ctx->upload_chunks - 1;
    ctx->upload_chunks = 0;
  }
  else
    launch_deskew_var_init(ctx->stream, ctx->d_scan, ctx->n_scan, ctx->d_poses, ctx->d_status, ctx->pv[0], var_init_params(ctx),
                           ctx->d_cache, 0, ctx->n_scan);
  ctx->n_pv[0] = ctx->n_scan;
  ctx->cache_is_reset = true;
  ctx->launches += 1;
  return mark_scan_read(ctx);
}

// The front of the per-scan step as two launches: k_deskew_var_init_down (vina_deskew + vina_var_init(ctx, 0) + the
// cache reset + the accumulation pass of vina_downsample) and k_down_emit_all (the rest of vina_downsample +
// vina_var_init(ctx, 1)). The count is published through mapped memory (resolve_n_down polls it).
int vn_front_fused(vina_ctx* ctx, const vina_imu_pose* poses, int m, const double R_end[9], const double p_end[3])
{
  if (m > VINA_MAX_POSES) return vn_fail(ctx, VINA_E_CAPACITY, "%d IMU poses > VINA_MAX_POSES", m);
  if (ctx->poses_in_flight) CU(cudaEventSynchronize(ctx->ev_poses));
  DeskewPoses* P = ctx->h_poses;
  P->m = m;
  memcpy(P->pose, poses, (size_t)m * sizeof(vina_imu_pose));
  memcpy(P->R_end, R_end, 72);
  memcpy(P->p_end, p_end, 24);
  memcpy(P->ext_R, ctx->cfg.ext_R, 72);
  memcpy(P->ext_t, ctx->cfg.ext_t, 24);
  CU(cudaMemcpyAsync(ctx->d_poses, P, sizeof(DeskewPoses), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaEventRecord(ctx->ev_poses, ctx->stream));
  ctx->poses_in_flight = true;
  const VarInitParams prm = var_init_params(ctx);
  launch_deskew_var_init_down(ctx->stream, ctx->d_scan, ctx->n_scan, ctx->d_poses, ctx->d_status, ctx->pv[0], prm, ctx->d_cache,
                              ctx->cfg.down_size, ctx->d_dtab, ctx->dmask, ctx->d_slot_of);
  ctx->n_pv[0] = ctx->n_scan;
  ctx->cache_is_reset = true;
  DownEmit de;
  de.n = ctx->n_scan;
  de.chunk = 0;
  de.tab = ctx->d_dtab;
  de.slot_of = ctx->d_slot_of;
  de.out = ctx->d_down;
  de.n_out_dev = ctx->d_n_down;
  de.pv = ctx->pv[1];
  de.prm = prm;
  de.counts = ctx->d_emit_counts;
  de.bar = ctx->d_emit_bar;
  de.pub = ctx->d_down_pub;
  de.seq = ++ctx->down_seq;
  de.status = ctx->d_status;
  int e = launch_down_emit_all(ctx->stream, de, ctx->sm_count);
  if (e) return vn_check_cuda(ctx, (cudaError_t)e, "k_down_emit_all launch");
  ctx->launches += 2;
  ctx->n_down_pending = true;
  ctx->n_down_mapped = true;
  return mark_scan_read(ctx);
}

extern "C" int vina_var_init(vina_ctx* ctx, int which)
{
  if (!ctx || which < 0 || which > 1) return VINA_E_ARG;
  VarInitParams prm = var_init_params(ctx);
  if (which == 1)
  {
    int r = resolve_n_down(ctx);
    if (r) return r;
  }
  const int n = which == 0 ? ctx->n_scan : ctx->n_down;
  launch_var_init(ctx->stream, which == 0 ? ctx->d_scan : ctx->d_down, nullptr, n, ctx->pv[which], prm);
  ctx->n_pv[which] = n;
  ctx->launches += 1;
  return which == 0 ? mark_scan_read(ctx) : VINA_OK;
}

extern "C" int vina_pvec_upload(vina_ctx* ctx, int which, const double* pnt, const double* var, int n)
{
  if (!ctx || which < 0 || which > 1 || !pnt || !var || n < 0) return VINA_E_ARG;
  if (n > ctx->cap_points) return vn_fail(ctx, VINA_E_CAPACITY, "%d points > max_scan_points", n);
  std::vector<double> tmp((size_t)n);
  for (int k = 0; k < 3; k++)
  {
    for (int i = 0; i < n; i++) tmp[i] = pnt[3 * (size_t)i + k];
    CU(cudaMemcpy(ctx->pv[which].p[k], tmp.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
  }
  const int ui[6] = { 0, 0, 0, 1, 1, 2 }, uj[6] = { 0, 1, 2, 1, 2, 2 };
  for (int k = 0; k < 6; k++)
  {
    for (int i = 0; i < n; i++) tmp[i] = var[9 * (size_t)i + ui[k] + 3 * uj[k]];
    CU(cudaMemcpy(ctx->pv[which].v[k], tmp.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
  }
  ctx->n_pv[which] = n;
  if (which == 1)
  {
    ctx->n_down = n;
    ctx->n_down_pending = false;
  }
  else
    ctx->n_scan = n;
  return VINA_OK;
}

extern "C" int vina_pvec_download(vina_ctx* ctx, int which, double* pnt, double* var, int cap)
{
  if (!ctx || which < 0 || which > 1 || !pnt || !var) return VINA_E_ARG;
  const int n = ctx->n_pv[which];
  if (cap < n) return VINA_E_ARG;
  CU(cudaStreamSynchronize(ctx->stream));
  std::vector<double> tmp((size_t)n);
  for (int k = 0; k < 3; k++)
  {
    CU(cudaMemcpy(tmp.data(), ctx->pv[which].p[k], (size_t)n * 8, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) pnt[3 * (size_t)i + k] = tmp[i];
  }
  const int ui[6] = { 0, 0, 0, 1, 1, 2 }, uj[6] = { 0, 1, 2, 1, 2, 2 };
  for (int k = 0; k < 6; k++)
  {
    CU(cudaMemcpy(tmp.data(), ctx->pv[which].v[k], (size_t)n * 8, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++)
    {
      var[9 * (size_t)i + ui[k] + 3 * uj[k]] = tmp[i];
      var[9 * (size_t)i + uj[k] + 3 * ui[k]] = tmp[i];
    }
  }
  return n;
}

// ---------------------------------------------------------------------------
extern "C" int vina_iekf_begin(vina_ctx* ctx, int which, const double rot_var[9], const double tsl_var[9])
{
  if (!ctx || which < 0 || which > 1 || !rot_var || !tsl_var) return VINA_E_ARG;
  memcpy(ctx->rot_var, rot_var, 72);
  memcpy(ctx->tsl_var, tsl_var, 72);
  ctx->iekf_which = which;
  const int n = ctx->n_pv[which];
  // the prior blocks of this call live in the device iterate (odometry.cpp:105-106)
  CU(cudaMemcpyAsync(ctx->d_iekf->rot_var, ctx->rot_var, 72, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_iekf->tsl_var, ctx->tsl_var, 72, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemsetAsync(&ctx->d_iekf->iter, 0, 4 * sizeof(int), ctx->stream));
  launch_fill_int(ctx->stream, ctx->d_cache, -1, n);  // vector<OctoTree*> octos(psize, nullptr), odometry.cpp:79
  ctx->iekf_blocks = iekf_grid_blocks(n, ctx->sm_count);
  ctx->launches += 1;
  return VINA_OK;
}

static int ensure_debug(vina_ctx* ctx)
{
  if (ctx->dbg.keys) return VINA_OK;
  const size_t cap = ctx->cap_points;
  CU(dalloc(&ctx->dbg.keys, 3 * cap));
  CU(dalloc(&ctx->dbg.codes, cap));
  CU(dalloc(&ctx->dbg.flags, cap));
  CU(dalloc(&ctx->dbg.sigma, cap));
  // dalloc's cudaMemset runs on the legacy default stream, which the ctx stream (cudaStreamNonBlocking) does
  // not wait for: without this a kernel launched right away can be overtaken by the zero-fill of its output
  CU(cudaDeviceSynchronize());
  return VINA_OK;
}

void vn_iekf_fill_seq(vina_ctx* ctx, IekfSeq* q, bool debug)
{
  const int w = ctx->iekf_which;
  q->pv_base = ctx->pv[w].p[0];
  q->pv_stride = ctx->cap_points;
  q->n_ptr = nullptr;
  q->n_host = ctx->n_pv[w];
  q->hmask = ctx->map.hmask;
  q->cache = ctx->d_cache;
  q->slots = ctx->map.slots;
  q->hot = ctx->map.hot;
  q->cold = ctx->map.cold;
  q->dev = ctx->d_iekf;
  q->partials = ctx->d_partials;
  q->ticket = ctx->d_ticket;
  q->result = ctx->d_result;
  q->voxel_size = ctx->cfg.voxel_size;
  q->seq = ctx->iekf_seq;
  q->pub = ctx->d_pub;
  q->pub_flag = ctx->d_pub_flag;
  q->pub_seq = ctx->pub_seq;
  IekfDebug none = { nullptr, nullptr, nullptr, nullptr };
  q->dbg = debug ? ctx->dbg : none;
}

// low-level accumulate: the caller's (R, p) go into the device iterate, the sums come back through mapped memory
int vn_iekf_launch(vina_ctx* ctx, const double R[9], const double p[3], bool debug)
{
  if (ctx->iekf_which < 0) return vn_fail(ctx, VINA_E_STATE, "vina_iekf_accumulate before vina_iekf_begin");
  if (debug)
  {
    int r = ensure_debug(ctx);
    if (r) return r;
  }
  CU(cudaMemcpyAsync(ctx->d_iekf->R, R, 72, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_iekf->p, p, 24, cudaMemcpyHostToDevice, ctx->stream));
  ++ctx->iekf_seq;
  IekfBatch bt;
  bt.mode = VN_IEKF_PUBLISH;
  bt.variant = ctx->iekf_variant;
  vn_iekf_fill_seq(ctx, &bt.s[0], debug);
  int e = launch_iekf(ctx->stream, bt, 1, ctx->iekf_blocks, debug);
  if (e) return vn_check_cuda(ctx, (cudaError_t)e, "k_iekf launch");
  ctx->dbg_valid = debug;
  ctx->launches += 1;
  return VINA_OK;
}

int vn_iterate_publish(vina_ctx* ctx, cudaStream_t st)
{
  launch_publish_iterate(st, ctx->d_iekf, ctx->d_pub, ctx->d_pub_flag, ++ctx->pub_seq);
  ctx->launches += 1;
  return vn_check_cuda(ctx, cudaGetLastError(), "k_publish_iterate");
}

int vn_iterate_wait(vina_ctx* ctx, cudaStream_t st)
{
  // `st` = the stream that carries the launch which hands the iterate over (the batch stream in batch replay)
  volatile unsigned long long* flag = ctx->h_pub_flag;
  const unsigned long long want = ctx->pub_seq;
  for (long spins = 0; *flag != want; spins++)
  {
    if ((spins & 0xfff) == 0xfff)
    {
      // a kernel may have failed, or the hand-over never happened: fall back to the stream state
      cudaError_t e = cudaStreamQuery(st);
      if (e == cudaSuccess)
      {
        if (*flag == want) break;
        return vn_fail(ctx, VINA_E_CUDA, "the IEKF loop finished without handing the iterate over");
      }
      if (e != cudaErrorNotReady) return vn_check_cuda(ctx, e, "IEKF loop");
    }
  }
  __sync_synchronize();
  return VINA_OK;
}

int vn_iekf_wait(vina_ctx* ctx)
{
  volatile unsigned long long* flag = reinterpret_cast<volatile unsigned long long*>(ctx->h_result) + 40;
  const unsigned long long want = ctx->iekf_seq;
  for (long spins = 0; *flag != want; spins++)
  {
    if ((spins & 0xfff) == 0xfff)
    {
      // the kernel may have failed: fall back to the stream state
      cudaError_t e = cudaStreamQuery(ctx->stream);
      if (e == cudaSuccess)
      {
        if (*flag == want) break;
        return vn_fail(ctx, VINA_E_CUDA, "k_iekf finished without publishing its result");
      }
      if (e != cudaErrorNotReady) return vn_check_cuda(ctx, e, "k_iekf");
    }
  }
  __sync_synchronize();
  return VINA_OK;
}

// expand the 34 packed sums
void vn_iekf_unpack(const double* r, double HTH[36], double HTz[6], double nnt[9], int32_t* match_num)
{
  int t = 0;
  for (int a = 0; a < 6; a++)
    for (int b = a; b < 6; b++, t++)
    {
      HTH[a + 6 * b] = r[t];
      HTH[b + 6 * a] = r[t];
    }
  for (int a = 0; a < 6; a++) HTz[a] = r[21 + a];
  const int ui[6] = { 0, 0, 0, 1, 1, 2 }, uj[6] = { 0, 1, 2, 1, 2, 2 };
  for (int k = 0; k < 6; k++)
  {
    nnt[ui[k] + 3 * uj[k]] = r[27 + k];
    nnt[uj[k] + 3 * ui[k]] = r[27 + k];
  }
  *match_num = (int32_t)(r[33] + 0.5);
}

extern "C" int vina_iekf_accumulate(vina_ctx* ctx, const double R[9], const double p[3], double HTH[36],
                                    double HTz[6], double nnt[9], int32_t* match_num)
{
  if (!ctx || !R || !p || !HTH || !HTz || !nnt || !match_num) return VINA_E_ARG;
  int r = vn_iekf_launch(ctx, R, p, false);
  if (r) return r;
  r = vn_iekf_wait(ctx);
  if (r) return r;
  vn_iekf_unpack(ctx->h_result, HTH, HTz, nnt, match_num);
  return VINA_OK;
}

extern "C" int vina_iekf_debug_assoc(vina_ctx* ctx, int64_t* keys, int32_t* codes, uint8_t* flags, double* sigma,
                                     int cap)
{
  if (!ctx || ctx->iekf_which < 0) return VINA_E_ARG;
  const int n = ctx->n_pv[ctx->iekf_which];
  if (cap < n) return VINA_E_ARG;
  if (!ctx->dbg_valid) return vn_fail(ctx, VINA_E_STATE, "no debug association recorded (use vina_iekf_accumulate_debug)");
  CU(cudaStreamSynchronize(ctx->stream));
  if (keys) CU(cudaMemcpy(keys, ctx->dbg.keys, (size_t)n * 24, cudaMemcpyDeviceToHost));
  if (codes) CU(cudaMemcpy(codes, ctx->dbg.codes, (size_t)n * 4, cudaMemcpyDeviceToHost));
  if (flags) CU(cudaMemcpy(flags, ctx->dbg.flags, (size_t)n, cudaMemcpyDeviceToHost));
  if (sigma) CU(cudaMemcpy(sigma, ctx->dbg.sigma, (size_t)n * 8, cudaMemcpyDeviceToHost));
  return n;
}

// same as vina_iekf_accumulate but also records the per-point association for vina_iekf_debug_assoc
extern "C" int vina_iekf_accumulate_debug(vina_ctx* ctx, const double R[9], const double p[3], double HTH[36],
                                          double HTz[6], double nnt[9], int32_t* match_num)
{
  if (!ctx || !R || !p || !HTH || !HTz || !nnt || !match_num) return VINA_E_ARG;
  int r = vn_iekf_launch(ctx, R, p, true);
  if (r) return r;
  r = vn_iekf_wait(ctx);
  if (r) return r;
  CU(cudaStreamSynchronize(ctx->stream));
  vn_iekf_unpack(ctx->h_result, HTH, HTz, nnt, match_num);
  return VINA_OK;
}

// ---------------------------------------------------------------------------
// start-up phase (host/vina_pipeline.cpp: vina_odom_init_scan)
int vn_init_ensure(vina_ctx* ctx)
{
  if (ctx->d_tree[0]) return VINA_OK;
  const size_t cap = ctx->cap_points;
  for (int k = 0; k < 2; k++) CU(dalloc(&ctx->d_tree[k], 2 * cap, false));
  CU(dalloc(&ctx->d_init_ds, cap));
  CU(dalloc(&ctx->d_init_dir, 3 * cap));
  CU(dalloc(&ctx->d_init_part, (cap / 128 + 2) * 28));
  CU(cudaHostAlloc((void**)&ctx->h_init_sums, 32 * sizeof(double), cudaHostAllocMapped));
  CU(cudaHostGetDevicePointer((void**)&ctx->d_init_sums, ctx->h_init_sums, 0));
  return VINA_OK;
}

int vn_map_clear(vina_ctx* ctx)
{
  MapView& M = ctx->map;
  CU(cudaStreamSynchronize(ctx->stream));
  int nn = 0;
  CU(cudaMemcpy(&nn, M.node_count, 4, cudaMemcpyDeviceToHost));
  if (nn > M.max_nodes) nn = M.max_nodes;
  // used records back to zero (= never-used pool memory, which the allocators rely on), table emptied, cursors reset
  if (nn > 0)
  {
    CU(cudaMemsetAsync(M.hot, 0, (size_t)nn * sizeof(NodeHot), ctx->stream));
    CU(cudaMemsetAsync(M.cold, 0, (size_t)nn * sizeof(NodeCold), ctx->stream));
  }
  launch_map_init(ctx->stream, M, ctx->hash_slots);
  CU(cudaMemsetAsync(M.node_count, 0, 4, ctx->stream));
  CU(cudaMemsetAsync(M.root_count, 0, 4, ctx->stream));
  CU(cudaMemsetAsync(M.win_cursor, 0, VINA_MAX_WIN * sizeof(int), ctx->stream));
  CU(cudaMemsetAsync(M.fix_cursor, 0, 4, ctx->stream));
  CU(cudaMemsetAsync(M.fixseg_cursor, 0, 4, ctx->stream));
  CU(cudaMemsetAsync(M.free_count, 0, 4, ctx->stream));
  CU(cudaMemsetAsync(M.free_seg_count, 0, 4, ctx->stream));
  CU(cudaMemsetAsync(M.slide_count, 0, 2 * sizeof(int), ctx->stream));
  M.slide_cur = 0;
  for (int i = 0; i < VINA_MAX_WIN; i++) M.mp[i] = i;
  ctx->ba_n = -1;
  return VINA_OK;
}

int vn_downsample_cloud(vina_ctx* ctx, const float4* in, int n, double size, float4* out, int* n_out)
{
  *n_out = 0;
  if (n <= 0) return VINA_OK;
  if (n > ctx->cap_points) return vn_fail(ctx, VINA_E_CAPACITY, "cloud of %d points > max_scan_points %d", n, ctx->cap_points);
  int k = launch_downsample(ctx->stream, in, n, size, ctx->d_dtab, ctx->dmask, ctx->d_slot_of, ctx->d_flag, ctx->d_scanbuf,
                            ctx->d_block_sums, ctx->d_n_down, out, ctx->d_status);
  if (k < 0) return vn_fail(ctx, VINA_E_CAPACITY, "cloud too large for the down-sampling scan kernels");
  ctx->launches += k;
  CU(cudaMemcpyAsync(ctx->h_n_down, ctx->d_n_down, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  *n_out = *ctx->h_n_down;
  return vn_check_status(ctx);
}

// one iteration of the kd-tree IEKF's point loop on the down-sampled pointVar set (pv[1]) against the local map
int vn_init_assoc(vina_ctx* ctx, const double R[9], const double p[3], int refind, double sums28[28])
{
  PoseD x;
  memcpy(x.R, R, 72);
  memcpy(x.p, p, 24);
  const int n = ctx->n_pv[1];
  for (int k = 0; k < 28; k++) ctx->h_init_sums[k] = 0.0;
  ctx->launches += launch_init_assoc(ctx->stream, ctx->pv[1], n, x, ctx->d_tree[ctx->tree_cur], ctx->n_tree, refind,
                                     ctx->d_init_ds, ctx->d_init_dir, ctx->d_init_part, ctx->d_init_sums);
  CU(cudaStreamSynchronize(ctx->stream));
  memcpy(sums28, ctx->h_init_sums, 28 * sizeof(double));
  return vn_check_cuda(ctx, cudaGetLastError(), "k_init_assoc");
}

// pl_tree += R pnt + p for the scan's points (no down-sampling here)
int vn_init_tree_push(vina_ctx* ctx, const double R[9], const double p[3])
{
  PoseD x;
  memcpy(x.R, R, 72);
  memcpy(x.p, p, 24);
  const int n = ctx->n_pv[1];
  if ((size_t)ctx->n_tree + n > 2 * (size_t)ctx->cap_points)
    return vn_fail(ctx, VINA_E_CAPACITY, "local map of the start-up phase: %d + %d points", ctx->n_tree, n);
  ctx->launches += launch_init_tree_append(ctx->stream, ctx->pv[1], n, x, ctx->d_tree[ctx->tree_cur] + ctx->n_tree);
  ctx->n_tree += n;
  return vn_check_cuda(ctx, cudaGetLastError(), "k_init_tree_append");
}

int vn_init_insert_frame(vina_ctx* ctx, const float* xyzt, int n, int n_skip, const vina_imu_pose* poses, int m,
                         const vina_state* x, int converged, int frame)
{
  if (n > ctx->cap_points) return vn_fail(ctx, VINA_E_CAPACITY, "frame of %d points > max_scan_points", n);
  if (m > VINA_MAX_POSES) return vn_fail(ctx, VINA_E_CAPACITY, "%d IMU poses > VINA_MAX_POSES", m);
  int r = ensure_front(ctx);  // (d_raw: the upload buffer)
  if (r) return r;
  CU(cudaStreamSynchronize(ctx->stream));  // the staging buffer of the pose table is reused
  DeskewPoses* hp = ctx->h_poses;
  hp->m = m;
  memcpy(hp->pose, poses, (size_t)m * sizeof(vina_imu_pose));
  memcpy(hp->R_end, x->R, 72);
  memcpy(hp->p_end, x->p, 24);
  memcpy(hp->ext_R, ctx->cfg.ext_R, 72);
  memcpy(hp->ext_t, ctx->cfg.ext_t, 24);
  CU(cudaMemcpyAsync(ctx->d_poses, hp, sizeof(DeskewPoses), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_raw, xyzt, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream));
  PoseD px;
  memcpy(px.R, x->R, 72);
  memcpy(px.p, x->p, 24);
  double rv[9], tv[9];
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < 3; i++)
    {
      rv[i + 3 * j] = x->cov[i + 15 * j];
      tv[i + 3 * j] = x->cov[(3 + i) + 15 * (3 + j)];
    }
  const VarInitParams prm = var_init_params(ctx);
  ctx->launches += launch_init_redeskew(ctx->stream, ctx->d_raw, n, n_skip, ctx->d_poses, px, rv, tv, converged, prm,
                                        ctx->pv[1], ctx->ins);
  // point 0 is pushed once more for every pose behind its own (see k_init_redeskew)
  int n_dup = 0;
  if (n_skip == 0 && n > 0)
  {
    int k0 = 0;
    while (k0 < m && !(poses[k0].t < (double)xyzt[3])) k0++;
    n_dup = m - 1 - k0 > 0 ? m - 1 - k0 : 0;
  }
  const int n_out = (n - n_skip) + n_dup;
  ctx->n_pv[1] = n_out;
  ctx->n_down = n_out;
  ctx->n_down_pending = false;
  if (n_out <= 0) return VINA_OK;
  // cut_voxel (voxel_map.cpp:4-45, the serial variant motion_init uses): no "fewer roots than threads" early-out
  MapView M = ctx->map;
  M.thread_num = 0;
  ctx->ins.stamp++;
  PoseD z;
  memset(&z, 0, sizeof(z));
  const double z9[9] = { 0 };
  ctx->launches += launch_map_insert_roots(ctx->stream, M, ctx->pv[1], nullptr, n_out, ctx->ins, z, z9, z9, 1);
  ctx->launches += launch_map_insert_leaves(ctx->stream, M, ctx->pv[1], nullptr, n_out, ctx->ins, frame);
  return vn_check_cuda(ctx, cudaGetLastError(), "start-up insert");
}

// ---------------------------------------------------------------------------
extern "C" int vina_map_insert(vina_ctx* ctx, int win_ord, const double R[9], const double p[3],
                               const double cov_rot[9], const double cov_tsl[9])
{
  if (!ctx || win_ord < 0 || win_ord >= ctx->cfg.win_size || !R || !p || !cov_rot || !cov_tsl) return VINA_E_ARG;
  PoseD x;
  memcpy(x.R, R, 72);
  memcpy(x.p, p, 24);
  ctx->ins.stamp++;
  const int n = ctx->n_pv[1];
  ctx->launches += launch_map_insert(ctx->stream, ctx->map, ctx->pv[1], nullptr, n, ctx->ins, win_ord, x, cov_rot, cov_tsl);
  return VINA_OK;
}

extern "C" int vina_map_recut(vina_ctx* ctx, int win_count, const vina_pose* x_buf)
{
  if (!ctx || win_count < 1 || win_count > ctx->cfg.win_size || !x_buf) return VINA_E_ARG;
  ctx->launches += launch_map_recut(ctx->stream, ctx->map, ctx->layers, win_count, reinterpret_cast<const PoseD*>(x_buf));
  return VINA_OK;
}

extern "C" int vina_map_margi(vina_ctx* ctx, int win_count, const vina_pose* x_buf)
{
  if (!ctx || win_count < 1 || win_count > ctx->cfg.win_size || !x_buf) return VINA_E_ARG;
  ctx->launches += launch_map_margi(ctx->stream, ctx->map, ctx->layers, win_count, reinterpret_cast<const PoseD*>(x_buf));
  ctx->map.slide_cur = 1 - ctx->map.slide_cur;
  return VINA_OK;
}

// the map update of the scan whose IEKF loop is still in flight: pose of the newest frame and posterior
// covariance blocks come from the device iterate (x_buf[win_count - 1] is ignored by the kernels)
int vn_map_insert_live(vina_ctx* ctx, int win_ord)
{
  PoseD x;
  memset(&x, 0, sizeof(x));
  const double z9[9] = { 0 };
  ctx->ins.stamp++;
  const int n = ctx->n_pv[1];
  // the node lists of the multi_recut that follows are collected on the side stream next to this insert's accumulation
  // (VINA_EARLY_COLLECT=0: in stream order, as part of the recut)
  EarlyCollect ec = { &ctx->layers, ctx->side_stream, ctx->ev_collect_fork, ctx->ev_collect_done };
  const bool early = ctx->early_collect && !ctx->split_overlap && ctx->side_stream && n > 0;
  ctx->launches += launch_map_insert(ctx->stream, ctx->map, ctx->pv[1], nullptr, n, ctx->ins, win_ord, x, z9, z9, ctx->d_iekf,
                                     early ? &ec : nullptr);
  ctx->collected_early = early;
  return VINA_OK;
}
int vn_map_recut_live(vina_ctx* ctx, int win_count, const vina_pose* x_buf)
{
  EarlyCollect ec = { &ctx->layers, ctx->side_stream, ctx->ev_collect_fork, ctx->ev_collect_done };
  ctx->launches += launch_map_recut(ctx->stream, ctx->map, ctx->layers, win_count, reinterpret_cast<const PoseD*>(x_buf),
                                    ctx->d_iekf, ctx->collected_early ? &ec : nullptr);
  ctx->collected_early = false;
  return VINA_OK;
}
int vn_map_margi_live(vina_ctx* ctx, int win_count, const vina_pose* x_buf)
{
  ctx->launches += launch_map_margi(ctx->stream, ctx->map, ctx->layers, win_count, reinterpret_cast<const PoseD*>(x_buf),
                                    ctx->d_iekf);
  ctx->map.slide_cur = 1 - ctx->map.slide_cur;
  return VINA_OK;
}

// multi_recut + multi_margi with the subdivisions next to the marginalisation (launch_map_recut_margi)
int vn_map_recut_margi_live(vina_ctx* ctx, int win_count, const vina_pose* x_buf)
{
  ctx->launches += launch_map_recut_margi(ctx->stream, ctx->side_stream, ctx->ev_fork, ctx->ev_join, ctx->map, ctx->layers,
                                          win_count, reinterpret_cast<const PoseD*>(x_buf), ctx->d_iekf);
  ctx->map.slide_cur = 1 - ctx->map.slide_cur;
  return vn_check_cuda(ctx, cudaGetLastError(), "recut + margi");
}

// sizes of the last map update (bench.py: algorithmic bytes of the map stages): [0] points inserted, [1] leaves they
// touched, [2..5] nodes under surf_map_slide per layer as multi_recut listed them, [6] leaves subdivided
extern "C" int vina_map_last_counts(vina_ctx* ctx, int32_t out[8])
{
  if (!ctx || !out) return VINA_E_ARG;
  CU(cudaStreamSynchronize(ctx->stream));
  int c[4] = { 0, 0, 0, 0 }, l[8] = { 0 }, sc[2] = { 0, 0 };
  CU(cudaMemcpy(c, ctx->ins.counters, sizeof(c), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(l, ctx->layers.count, sizeof(l), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(sc, ctx->map.slide_count, sizeof(sc), cudaMemcpyDeviceToHost));
  out[0] = ctx->n_pv[1];
  out[1] = c[1];
  out[2] = sc[ctx->map.slide_cur];
  out[3] = l[1];
  out[4] = l[2];
  out[5] = l[3];
  out[6] = l[4];
  out[7] = 0;
  return VINA_OK;
}

extern "C" int vina_map_shift_window(vina_ctx* ctx)
{
  if (!ctx) return VINA_E_ARG;
  // the slot of the marginalised frame is free again: reset its arena
  const int freed = ctx->map.mp[0];
  CU(cudaMemsetAsync(ctx->map.win_cursor + freed, 0, sizeof(int), ctx->stream));
  for (int i = 0; i < ctx->cfg.win_size; i++)  // local_mapping.cpp:521-526, mgsize = 1
  {
    ctx->map.mp[i] += 1;
    if (ctx->map.mp[i] >= ctx->cfg.win_size) ctx->map.mp[i] -= ctx->cfg.win_size;
  }
  return VINA_OK;
}

extern "C" int64_t vina_map_count(vina_ctx* ctx, int64_t* n_roots, int64_t* n_slide)
{
  if (!ctx) return VINA_E_ARG;
  int r = vn_check_status(ctx);
  if (r) return r;
  int nn = 0, nfree = 0, sc[2] = { 0, 0 };
  CU(cudaMemcpy(&nn, ctx->map.node_count, 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(&nfree, ctx->map.free_count, 4, cudaMemcpyDeviceToHost));
  if (nfree > 0) nn -= nfree;  // ids on the free stack of the map pruning
  CU(cudaMemcpy(sc, ctx->map.slide_count, 8, cudaMemcpyDeviceToHost));
  if (n_slide) *n_slide = sc[ctx->map.slide_cur];
  if (n_roots)
  {
    int nr = 0;
    CU(cudaMemcpy(&nr, ctx->map.root_count, 4, cudaMemcpyDeviceToHost));
    *n_roots = nr;
  }
  return nn;
}

extern "C" int64_t vina_map_export(vina_ctx* ctx, vina_node_record* out, int64_t cap)
{
  if (!ctx || !out) return VINA_E_ARG;
  int r = vn_check_status(ctx);
  if (r) return r;
  int nn = 0;
  CU(cudaMemcpy(&nn, ctx->map.node_count, 4, cudaMemcpyDeviceToHost));
  if (nn > ctx->map.max_nodes) nn = ctx->map.max_nodes;
  if (nn == 0) return 0;
  int nfree = 0;
  CU(cudaMemcpy(&nfree, ctx->map.free_count, 4, cudaMemcpyDeviceToHost));
  if (nfree < 0) nfree = 0;
  if (cap < nn - nfree) return VINA_E_ARG;
  vina_node_record* d_out = nullptr;
  long long* d_cnt = nullptr;
  CU(cudaMalloc((void**)&d_out, (size_t)nn * sizeof(vina_node_record)));
  CU(cudaMalloc((void**)&d_cnt, 8));
  launch_map_export(ctx->stream, ctx->map, d_out, nn, d_cnt);
  CU(cudaStreamSynchronize(ctx->stream));
  if (nfree == 0)
    CU(cudaMemcpy(out, d_out, (size_t)nn * sizeof(vina_node_record), cudaMemcpyDeviceToHost));
  else
  {
    // records on the free stack (layer == -1) are dropped
    std::vector<vina_node_record> all((size_t)nn);
    CU(cudaMemcpy(all.data(), d_out, (size_t)nn * sizeof(vina_node_record), cudaMemcpyDeviceToHost));
    int64_t k = 0;
    for (int i = 0; i < nn; i++)
      if (all[i].layer >= 0)
      {
        if (k >= cap)
        {
          cudaFree(d_out);
          cudaFree(d_cnt);
          return VINA_E_ARG;
        }
        out[k++] = all[i];
      }
    nn = (int)k;
  }
  cudaFree(d_out);
  cudaFree(d_cnt);
  return nn;
}

// ---------------------------------------------------------------------------
// map pruning: the `else if (release_flag)` branch of the idle path (local_mapping.cpp:317-341)
extern "C" int vina_map_set_journey(vina_ctx* ctx, double jour)
{
  if (!ctx) return VINA_E_ARG;
  ctx->map.jour = jour;
  return VINA_OK;
}

extern "C" int vina_map_prune(vina_ctx* ctx, double jour, int horizon, int64_t* roots_erased, int64_t* nodes_freed)
{
  if (!ctx) return VINA_E_ARG;
  if (roots_erased) *roots_erased = 0;
  if (nodes_freed) *nodes_freed = 0;
  if (horizon <= 0) horizon = 700;
  launch_map_prune_mark(ctx->stream, ctx->map, jour, horizon, ctx->d_prune);
  ctx->launches += 1;
  int cnt[4] = { 0, 0, 0, 0 }, cursor = 0;
  CU(cudaMemcpyAsync(cnt, ctx->d_prune, sizeof(cnt), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(&cursor, ctx->map.fix_cursor, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  int r = vn_check_status(ctx);
  if (r) return r;
  if (cnt[0] == 0) return VINA_OK;  // nothing is stale: the map is left exactly as it is
  // scratch for the compaction of the fixed-point pool (an idle-time operation: a transient allocation of at
  // most the pool's used part; without it the nodes and the hash slots are still given back)
  PointRec* d_tmp = nullptr;
  if (cursor > 0 && cudaMalloc((void**)&d_tmp, (size_t)cursor * sizeof(PointRec)) != cudaSuccess)
  {
    cudaGetLastError();
    d_tmp = nullptr;
  }
  launch_map_prune_sweep(ctx->stream, ctx->map, ctx->hash_slots, d_tmp, ctx->d_prune);
  ctx->launches += 3;
  CU(cudaMemcpyAsync(cnt, ctx->d_prune, sizeof(cnt), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (d_tmp && cnt[2] > 0)
    CU(cudaMemcpyAsync(ctx->map.fix_pool, d_tmp, (size_t)cnt[2] * sizeof(PointRec), cudaMemcpyDeviceToDevice,
                       ctx->stream));
  launch_map_prune_finish(ctx->stream, ctx->map, ctx->d_prune, d_tmp ? 1 : 0);
  ctx->launches += 1;
  CU(cudaStreamSynchronize(ctx->stream));
  if (d_tmp) cudaFree(d_tmp);
  if (roots_erased) *roots_erased = cnt[0];
  if (nodes_freed) *nodes_freed = cnt[1];
  return vn_check_status(ctx);
}

// ---------------------------------------------------------------------------
// map sharded by voxel-hash range (shard_kernels.cu)
extern "C" int vina_shard_owner(int64_t kx, int64_t ky, int64_t kz, int world)
{
  unsigned long long key;
  if (world < 1 || world > VINA_MAX_WORLD || !pack_key(kx, ky, kz, &key)) return -1;
  return shard_owner(key, world);
}

static int ensure_shard(vina_ctx* ctx)
{
  if (ctx->d_sh_owner) return VINA_OK;
  const size_t cap = ctx->cap_points;
  CU(dalloc(&ctx->d_sh_owner, cap));
  CU(dalloc(&ctx->d_sh_hist, (cap / 256 + 2) * VINA_MAX_WORLD));
  CU(dalloc(&ctx->d_sh_counts, 2 * VINA_MAX_WORLD + 2));
  CU(cudaHostAlloc((void**)&ctx->h_sh_counts, (2 * VINA_MAX_WORLD + 2) * sizeof(int), cudaHostAllocDefault));
  CU(cudaDeviceSynchronize());  // see ensure_debug
  return VINA_OK;
}

extern "C" int vina_shard_route(vina_ctx* ctx, int world, int first, int count, int64_t scan_index_base,
                                const double R[9], const double p[3], const double cov_rot[9],
                                const double cov_tsl[9], void* d_send, int32_t* counts_out)
{
  if (!ctx || world < 1 || world > VINA_MAX_WORLD || first < 0 || count < 0 || !R || !p || !cov_rot || !cov_tsl ||
      !counts_out || (count > 0 && !d_send))
    return VINA_E_ARG;
  if (first + count > ctx->n_pv[1])
    return vn_fail(ctx, VINA_E_ARG, "slice [%d, %d) exceeds the %d down-sampled points", first, first + count,
                   ctx->n_pv[1]);
  int r = ensure_shard(ctx);
  if (r) return r;
  PoseD x;
  memcpy(x.R, R, 72);
  memcpy(x.p, p, 24);
  ctx->launches += launch_shard_route(ctx->stream, ctx->pv[1], first, count, x, cov_rot, cov_tsl, ctx->cfg.voxel_size,
                                      world, ctx->d_sh_owner, ctx->d_sh_hist, ctx->d_sh_counts,
                                      ctx->d_sh_counts + VINA_MAX_WORLD, (double*)d_send, (long long)scan_index_base,
                                      ctx->d_status, false);
  CU(cudaMemcpyAsync(ctx->h_sh_counts, ctx->d_sh_counts, world * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < world; k++) counts_out[k] = ctx->h_sh_counts[k];
  return VINA_OK;
}

// ---- the exchange fused into the routing kernel (peer stores over NVLink) -----------------------------
extern "C" int vina_shard_p2p_create(vina_ctx* ctx, int rank, int world, int64_t inbox_records, void* ipc_handles_out)
{
  if (!ctx || world < 1 || world > VINA_MAX_WORLD || rank < 0 || rank >= world || inbox_records < 1) return VINA_E_ARG;
  if (ctx->p2p_inbox) return vn_fail(ctx, VINA_E_STATE, "vina_shard_p2p_create called twice");
  int r = ensure_shard(ctx);
  if (r) return r;
  if (inbox_records > ctx->cap_points) inbox_records = ctx->cap_points;  // the insert works on max_scan_points at most
  // two regions: map-build records (13 doubles each), then association queries (10 doubles each)
  CU(dalloc(&ctx->p2p_inbox, (size_t)inbox_records * (VINA_SHARD_RECORD_DOUBLES + VINA_SHARD_QUERY_DOUBLES), false));
  CU(dalloc(&ctx->p2p_ctrl, 1));
  CU(dalloc(&ctx->d_n_recv, 1));
  CU(dalloc(&ctx->d_p2p_base, VINA_MAX_WORLD));
  CU(cudaDeviceSynchronize());
  ctx->p2p_cap = inbox_records;
  ctx->peers.rank = rank;
  ctx->peers.world = world;
  ctx->peers.inbox[rank] = ctx->p2p_inbox;
  ctx->peers.ctrl[rank] = ctx->p2p_ctrl;
  if (ipc_handles_out)
  {
    cudaIpcMemHandle_t h[2];
    CU(cudaIpcGetMemHandle(&h[0], ctx->p2p_inbox));
    CU(cudaIpcGetMemHandle(&h[1], ctx->p2p_ctrl));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles are exchanged as 64-byte blobs");
    memcpy(ipc_handles_out, h, sizeof(h));
  }
  ctx->p2p_connected = world == 1;
  return VINA_OK;
}

extern "C" int vina_shard_p2p_connect(vina_ctx* ctx, const void* all_handles)
{
  if (!ctx || !all_handles || !ctx->p2p_inbox) return VINA_E_ARG;
  const cudaIpcMemHandle_t* h = reinterpret_cast<const cudaIpcMemHandle_t*>(all_handles);
  for (int q = 0; q < ctx->peers.world; q++)
  {
    if (q == ctx->peers.rank) continue;
    void *pi = nullptr, *pc = nullptr;
    CU(cudaIpcOpenMemHandle(&pi, h[2 * q + 0], cudaIpcMemLazyEnablePeerAccess));
    CU(cudaIpcOpenMemHandle(&pc, h[2 * q + 1], cudaIpcMemLazyEnablePeerAccess));
    ctx->p2p_opened[2 * q + 0] = pi;
    ctx->p2p_opened[2 * q + 1] = pc;
    ctx->peers.inbox[q] = (double*)pi;
    ctx->peers.ctrl[q] = (ShardCtrl*)pc;
  }
  ctx->p2p_connected = true;
  return VINA_OK;
}

extern "C" int vina_shard_p2p_pointers(vina_ctx* ctx, void** inbox, void** ctrl)
{
  if (!ctx || !inbox || !ctrl || !ctx->p2p_inbox) return VINA_E_ARG;
  *inbox = ctx->p2p_inbox;
  *ctrl = ctx->p2p_ctrl;
  return VINA_OK;
}

extern "C" int vina_shard_p2p_connect_local(vina_ctx* ctx, void* const* inbox_ptrs, void* const* ctrl_ptrs)
{
  if (!ctx || !inbox_ptrs || !ctrl_ptrs || !ctx->p2p_inbox) return VINA_E_ARG;
  for (int q = 0; q < ctx->peers.world; q++)
  {
    if (!inbox_ptrs[q] || !ctrl_ptrs[q]) return VINA_E_ARG;
    ctx->peers.inbox[q] = (double*)inbox_ptrs[q];
    ctx->peers.ctrl[q] = (ShardCtrl*)ctrl_ptrs[q];
  }
  ctx->p2p_connected = true;
  return VINA_OK;
}

extern "C" int vina_shard_route_p2p(vina_ctx* ctx, int first, int count, int64_t scan_index_base, const double R[9],
                                    const double p[3], const double cov_rot[9], const double cov_tsl[9], int phase)
{
  if (!ctx || first < 0 || count < 0 || !R || !p || !cov_rot || !cov_tsl || phase < 0 || phase > 2) return VINA_E_ARG;
  if (!ctx->p2p_connected) return vn_fail(ctx, VINA_E_STATE, "vina_shard_route_p2p before the peers are connected");
  if (first + count > ctx->n_pv[1])
    return vn_fail(ctx, VINA_E_ARG, "slice [%d, %d) exceeds the %d down-sampled points", first, first + count,
                   ctx->n_pv[1]);
  PoseD x;
  memcpy(x.R, R, 72);
  memcpy(x.p, p, 24);
  if (phase != 2) ++ctx->p2p_epoch;
  ctx->launches += launch_shard_route_p2p(ctx->stream, ctx->pv[1], first, count, x, cov_rot, cov_tsl, ctx->cfg.voxel_size,
                                          ctx->peers, ctx->d_sh_owner, ctx->d_sh_hist, ctx->d_sh_counts,
                                          ctx->d_sh_counts + VINA_MAX_WORLD, ctx->d_p2p_base, ctx->p2p_epoch,
                                          (long long)scan_index_base, ctx->p2p_cap, ctx->d_status, phase);
  return vn_check_cuda(ctx, cudaGetLastError(), "p2p route");
}

extern "C" int vina_shard_insert_begin_p2p(vina_ctx* ctx, int win_ord, int32_t* n_recv, int32_t* local_roots,
                                           int32_t* local_slide)
{
  if (!ctx || win_ord < 0 || win_ord >= ctx->cfg.win_size || !n_recv || !local_roots || !local_slide) return VINA_E_ARG;
  if (!ctx->p2p_connected || ctx->p2p_epoch == 0) return vn_fail(ctx, VINA_E_STATE, "no routed scan to insert");
  const int cap = (int)ctx->p2p_cap;
  // wait for the peers' records (device side), unpack; the count stays on the device for the kernels below
  ctx->launches += launch_shard_recv_p2p(ctx->stream, ctx->peers, ctx->p2p_epoch, ctx->d_n_recv, cap, ctx->pv[1], ctx->ins,
                                         ctx->d_status);
  ctx->ins.stamp++;
  PoseD x;
  memset(&x, 0, sizeof(x));
  double z9[9] = { 0 };
  ctx->launches += launch_map_insert_roots(ctx->stream, ctx->map, ctx->pv[1], ctx->d_n_recv, cap, ctx->ins, x, z9, z9, 1);
  CU(cudaMemcpyAsync(ctx->h_sh_counts, ctx->ins.counters, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(ctx->h_sh_counts + 1, ctx->map.slide_count + ctx->map.slide_cur, sizeof(int),
                     cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(ctx->h_sh_counts + 2, ctx->d_n_recv, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  *local_roots = ctx->h_sh_counts[0];
  *local_slide = ctx->h_sh_counts[1];
  *n_recv = ctx->h_sh_counts[2];
  ctx->n_pv[1] = ctx->h_sh_counts[2];
  ctx->n_down = ctx->n_pv[1];
  ctx->n_down_pending = false;
  return vn_check_status(ctx);
}

extern "C" int vina_shard_query_route(vina_ctx* ctx, int world, int first, int count, int64_t scan_index_base,
                                      const double R[9], const double p[3], void* d_send, int32_t* counts_out)
{
  if (!ctx || world < 1 || world > VINA_MAX_WORLD || first < 0 || count < 0 || !R || !p || !counts_out ||
      (count > 0 && !d_send))
    return VINA_E_ARG;
  if (first + count > ctx->n_pv[0])
    return vn_fail(ctx, VINA_E_ARG, "slice [%d, %d) exceeds the %d scan points", first, first + count, ctx->n_pv[0]);
  int r = ensure_shard(ctx);
  if (r) return r;
  PoseD x;
  memcpy(x.R, R, 72);
  memcpy(x.p, p, 24);
  double z9[9] = { 0 };
  ctx->launches += launch_shard_route(ctx->stream, ctx->pv[0], first, count, x, z9, z9, ctx->cfg.voxel_size, world,
                                      ctx->d_sh_owner, ctx->d_sh_hist, ctx->d_sh_counts,
                                      ctx->d_sh_counts + VINA_MAX_WORLD, (double*)d_send, (long long)scan_index_base,
                                      ctx->d_status, true);
  CU(cudaMemcpyAsync(ctx->h_sh_counts, ctx->d_sh_counts, world * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < world; k++) counts_out[k] = ctx->h_sh_counts[k];
  return VINA_OK;
}

extern "C" int vina_shard_query_accumulate(vina_ctx* ctx, const void* d_recv, int n, const double R[9],
                                           const double p[3], const double rot_var[9], const double tsl_var[9],
                                           double* d_sums34)
{
  if (!ctx || n < 0 || (n > 0 && !d_recv) || !R || !p || !rot_var || !tsl_var || !d_sums34) return VINA_E_ARG;
  if (n > ctx->cap_points) return vn_fail(ctx, VINA_E_CAPACITY, "%d routed points > max_scan_points %d", n, ctx->cap_points);
  // the received records are evaluated from the query set pv[1] (the full-scan set pv[0] keeps the caller's
  // points for the next iteration's routing); no leaf cache: every point looks its voxel up
  ctx->launches += launch_shard_unpack_query(ctx->stream, (const double*)d_recv, n, ctx->pv[1]);
  ctx->n_pv[1] = n;
  ctx->n_down = n;
  ctx->n_down_pending = false;
  int r = vina_iekf_begin(ctx, 1, rot_var, tsl_var);
  if (r) return r;
  CU(cudaMemcpyAsync(ctx->d_iekf->R, R, 72, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_iekf->p, p, 24, cudaMemcpyHostToDevice, ctx->stream));
  IekfBatch bt;
  bt.mode = 0;  // sums stay in the device iterate
  bt.variant = 0;
  vn_iekf_fill_seq(ctx, &bt.s[0], false);
  int e = launch_iekf(ctx->stream, bt, 1, ctx->iekf_blocks, false);
  if (e) return vn_check_cuda(ctx, (cudaError_t)e, "k_iekf launch");
  ctx->launches += 1;
  CU(cudaMemcpyAsync(d_sums34, ctx->d_iekf->sums, VN_IEKF_NACC * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  return VINA_OK;
}

// The IEKF loop against the sharded map with everything on the device (see include/vina_b200.h): per iteration
// route + peer stores of the queries (query channel of the inboxes), k_iekf on what arrived, the 34 sums to every
// peer, rank-ordered total and update. The device iterate (ctx->d_iekf) must be staged on every rank alike.
int vn_shard_iekf_enqueue(vina_ctx* ctx, int first, int count, int max_iter, int part)
{
  if (!ctx->p2p_connected) return vn_fail(ctx, VINA_E_STATE, "sharded IEKF before the peers are connected");
  if (first < 0 || count < 0 || first + count > ctx->n_pv[0])
    return vn_fail(ctx, VINA_E_ARG, "slice [%d, %d) exceeds the %d scan points", first, first + count, ctx->n_pv[0]);
  const int cap = (int)ctx->p2p_cap;
  ctx->iekf_which = 1;  // the received queries are evaluated from the query set pv[1]
  ctx->n_pv[1] = 0;     // (its size lives on the device: d_n_recv)
  ctx->n_down = 0;
  ctx->n_down_pending = false;
  ctx->iekf_blocks = iekf_grid_blocks(cap, ctx->sm_count);
  ctx->dbg_valid = false;
  IekfBatch bt;
  bt.mode = VN_IEKF_GATED | VN_IEKF_NOCACHE;
  bt.variant = 0;
  vn_iekf_fill_seq(ctx, &bt.s[0], false);
  bt.s[0].n_ptr = ctx->d_n_recv;
  bt.s[0].n_host = 0;
  // part 0: max_iter whole iterations; 1..4: the four parts of ONE iteration (1 never waits; 2 waits for the
  // peers' part 1, 3 for their part 2, 4 for their part 3) - several ranks driven from one host thread
  const int iters = part == 0 ? max_iter : 1;
  for (int it = 0; it < iters; it++)
  {
    if (part <= 1) ++ctx->p2p_qepoch;
    const unsigned long long epoch = ctx->p2p_qepoch;
    for (int ph = 1; ph <= 3; ph++)
      if (part == 0 || part == ph)
        ctx->launches += launch_shard_query_p2p(ctx->stream, ctx->pv[0], first, count, ctx->d_iekf, ctx->cfg.voxel_size,
                                                ctx->peers, ctx->d_sh_owner, ctx->d_sh_hist, ctx->d_sh_counts,
                                                ctx->d_sh_counts + VINA_MAX_WORLD, ctx->d_p2p_base, epoch, ctx->p2p_cap,
                                                ctx->d_n_recv, ctx->pv[1], ctx->d_status, ph);
    if (part == 0 || part == 3)
    {
      int e = launch_iekf(ctx->stream, bt, 1, ctx->iekf_blocks, false);
      if (e) return vn_check_cuda(ctx, (cudaError_t)e, "k_iekf launch");
      launch_p2p_sums_publish(ctx->stream, ctx->peers, ctx->d_iekf, epoch);
      ctx->launches += 2;
    }
    if (part == 0 || part == 4)
    {
      launch_p2p_sums_solve(ctx->stream, ctx->peers, ctx->d_iekf, epoch, ctx->d_status);
      ctx->launches += 1;
    }
  }
  return vn_check_cuda(ctx, cudaGetLastError(), "sharded IEKF loop");
}

extern "C" int vina_shard_insert_begin(vina_ctx* ctx, const void* d_recv, int n, int win_ord, int32_t* local_roots,
                                       int32_t* local_slide)
{
  if (!ctx || n < 0 || (n > 0 && !d_recv) || win_ord < 0 || win_ord >= ctx->cfg.win_size || !local_roots || !local_slide)
    return VINA_E_ARG;
  if (n > ctx->cap_points) return vn_fail(ctx, VINA_E_CAPACITY, "%d routed points > max_scan_points %d", n, ctx->cap_points);
  int r = ensure_shard(ctx);
  if (r) return r;
  // the received records replace the down-sampled pointVar set of this ctx (body points; the covariance
  // arrays are not used: the world covariance arrives in the record)
  ctx->launches += launch_shard_unpack(ctx->stream, (const double*)d_recv, n, ctx->pv[1], ctx->ins);
  ctx->n_pv[1] = n;
  ctx->n_down = n;
  ctx->n_down_pending = false;
  ctx->ins.stamp++;
  PoseD x;
  memset(&x, 0, sizeof(x));
  double z9[9] = { 0 };
  ctx->launches += launch_map_insert_roots(ctx->stream, ctx->map, ctx->pv[1], nullptr, n, ctx->ins, x, z9, z9, 1);
  CU(cudaMemcpyAsync(ctx->h_sh_counts, ctx->ins.counters, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(ctx->h_sh_counts + 1, ctx->map.slide_count + ctx->map.slide_cur, sizeof(int),
                     cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  *local_roots = ctx->h_sh_counts[0];
  *local_slide = ctx->h_sh_counts[1];
  return VINA_OK;
}

extern "C" int vina_shard_insert_finish(vina_ctx* ctx, int win_ord, int global_roots, int global_slide)
{
  if (!ctx || win_ord < 0 || win_ord >= ctx->cfg.win_size || global_roots < 0 || global_slide < 0) return VINA_E_ARG;
  // the early-out of voxel_map.cpp:96-97 compares the scan's distinct roots over the WHOLE map
  launch_fill_int(ctx->stream, ctx->ins.counters, global_roots, 1);
  ctx->map.slide_others = global_slide - ctx->h_sh_counts[1];
  if (ctx->map.slide_others < 0) ctx->map.slide_others = 0;
  ctx->launches += 1 + launch_map_insert_leaves(ctx->stream, ctx->map, ctx->pv[1], nullptr, ctx->n_pv[1], ctx->ins, win_ord);
  return VINA_OK;
}

// ---------------------------------------------------------------------------
// BA LiDAR factor (ba_kernels.cu)
static int ensure_ba(vina_ctx* ctx)
{
  if (ctx->d_ba) return VINA_OK;
  long long cap = (long long)ctx->map.max_nodes;
  if (cap > (1 << 20)) cap = 1 << 20;  // 1.1 GB of factors at most
  CU(dalloc(&ctx->d_ba, (size_t)cap, false));
  CU(dalloc(&ctx->d_ba_n, 1));
  CU(dalloc(&ctx->d_ba_partial, ba_partial_doubles(ctx->sm_count), false));
  CU(dalloc(&ctx->d_ba_out, (size_t)36 * VINA_MAX_WIN * VINA_MAX_WIN + 6 * VINA_MAX_WIN + 8));
  CU(dalloc(&ctx->d_ba_lam, (size_t)cap, false));
  // results are written by the kernels straight into mapped pinned memory, completion is a polled sequence number
  CU(cudaHostAlloc((void**)&ctx->h_ba_out, ((size_t)36 * VINA_MAX_WIN * VINA_MAX_WIN + 6 * VINA_MAX_WIN + 8 + 4096) * sizeof(double),
                   cudaHostAllocMapped));
  CU(cudaHostGetDevicePointer((void**)&ctx->d_ba_map, ctx->h_ba_out, 0));
  CU(cudaHostAlloc((void**)&ctx->h_ba_flag, 64, cudaHostAllocMapped));
  CU(cudaHostGetDevicePointer((void**)&ctx->d_ba_flag, ctx->h_ba_flag, 0));
  *ctx->h_ba_flag = 0ull;
  CU(dalloc(&ctx->d_ba_ticket, 1));
  CU(cudaDeviceSynchronize());  // see ensure_debug
  ctx->ba_cap = (int)cap;
  return VINA_OK;
}

int vn_ba_collect_enqueue(vina_ctx* ctx)
{
  int r = ensure_ba(ctx);
  if (r) return r;
  ctx->launches += launch_ba_collect(ctx->stream, ctx->map, ctx->layers, ctx->d_ba, ctx->d_ba_n, ctx->ba_cap);
  ctx->ba_n = -1;
  return vn_check_cuda(ctx, cudaGetLastError(), "k_ba_collect");
}

int vn_ba_normal_scatter(vina_ctx* ctx, double nnt[9])
{
  for (int k = 0; k < 9; k++) nnt[k] = 0.0;
  int32_t n = 0;
  int r = vina_ba_count(ctx, &n);
  if (r) return r;
  if (n <= 0) return VINA_OK;
  // (a few thousand factors, once per motion_init convergence check: read the store back)
  std::vector<BaFactor> h((size_t)n);
  CU(cudaMemcpyAsync(h.data(), ctx->d_ba, (size_t)n * sizeof(BaFactor), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  for (int a = 0; a < n; a++)
  {
    const double* v = h[a].eig_vector;  // column 0
    for (int c = 0; c < 3; c++)
      for (int rr = 0; rr < 3; rr++) nnt[rr + 3 * c] += v[rr] * v[c];
  }
  return VINA_OK;
}

int vn_ba_writeback_enqueue(vina_ctx* ctx)
{
  if (!ctx->d_ba) return VINA_OK;
  ctx->launches += launch_ba_writeback(ctx->stream, ctx->map, ctx->d_ba, ctx->d_ba_n, ctx->sm_count);
  return vn_check_cuda(ctx, cudaGetLastError(), "k_ba_writeback");
}

extern "C" int vina_ba_set_capture(vina_ctx* ctx, int on)
{
  if (!ctx) return VINA_E_ARG;
  ctx->ba_capture = on != 0;
  return VINA_OK;
}

extern "C" int vina_ba_collect(vina_ctx* ctx, int32_t* n_factors)
{
  if (!ctx) return VINA_E_ARG;
  int r = vn_ba_collect_enqueue(ctx);
  if (r) return r;
  return vina_ba_count(ctx, n_factors);
}

extern "C" int vina_ba_count(vina_ctx* ctx, int32_t* n_factors)
{
  if (!ctx || !n_factors) return VINA_E_ARG;
  if (!ctx->d_ba) return vn_fail(ctx, VINA_E_STATE, "no BA factors collected yet");
  if (ctx->ba_n < 0)
  {
    int n = 0;
    CU(cudaMemcpyAsync(&n, ctx->d_ba_n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->ba_n = n < ctx->ba_cap ? n : ctx->ba_cap;
  }
  *n_factors = ctx->ba_n;
  return vn_check_status(ctx);
}

// wait for the completion flag of the last BA evaluation (mapped memory, polled; falls back to the stream state so
// that a failed kernel is reported instead of spinning forever)
static int ba_wait(vina_ctx* ctx, const char* what)
{
  volatile unsigned long long* flag = ctx->h_ba_flag;
  const unsigned long long want = ctx->ba_seq;
  for (long spins = 0; *flag != want; spins++)
    if ((spins & 0xfff) == 0xfff)
    {
      cudaError_t e = cudaStreamQuery(ctx->stream);
      if (e == cudaSuccess)
      {
        if (*flag == want) break;
        return vn_fail(ctx, VINA_E_CUDA, "%s finished without signalling", what);
      }
      if (e != cudaErrorNotReady) return vn_check_cuda(ctx, e, what);
    }
  __sync_synchronize();
  return VINA_OK;
}

// the two halves of vina_ba_lidar_hessian: the host can do its own work (the IMU factors) while the kernels run
int vn_ba_hess_enqueue(vina_ctx* ctx, const vina_pose* xs, int win)
{
  if (!ctx->d_ba) return vn_fail(ctx, VINA_E_STATE, "no BA factors collected yet");
  if (win != ctx->cfg.win_size) return vn_fail(ctx, VINA_E_ARG, "win %d != LocalBA.win_size %d", win, ctx->cfg.win_size);
  // (every entry of Hess / JacT / residual is written by k_ba_reduce: no clearing needed)
  const BaDone done = { ctx->d_ba_ticket, ctx->d_ba_flag, ++ctx->ba_seq };
  ctx->launches += launch_ba_hess(ctx->stream, ctx->d_ba, ctx->d_ba_n, reinterpret_cast<const PoseD*>(xs), win, ctx->sm_count,
                                  ctx->d_ba_partial, ctx->d_ba_map, done);
  return vn_check_cuda(ctx, cudaGetLastError(), "k_ba_hess");
}
int vn_ba_hess_finish(vina_ctx* ctx, int win, double* Hess, double* JacT, double* residual)
{
  const size_t dim = 6 * (size_t)win;
  int r = ba_wait(ctx, "k_ba_hess");
  if (r) return r;
  const double* h = ctx->h_ba_out;
  memcpy(Hess, h, dim * dim * sizeof(double));
  memcpy(JacT, h + dim * dim, dim * sizeof(double));
  *residual = h[dim * dim + dim];
  return VINA_OK;
}

extern "C" int vina_ba_lidar_hessian(vina_ctx* ctx, const vina_pose* xs, int win, double* Hess, double* JacT,
                                     double* residual)
{
  if (!ctx || !xs || win < 1 || win > VINA_MAX_WIN || !Hess || !JacT || !residual) return VINA_E_ARG;
  int r = vn_ba_hess_enqueue(ctx, xs, win);
  if (r) return r;
  return vn_ba_hess_finish(ctx, win, Hess, JacT, residual);
}

extern "C" int vina_ba_lidar_residual(vina_ctx* ctx, const vina_pose* xs, int win, double* residual, double* lam0, int cap)
{
  if (!ctx || !xs || win < 1 || win > VINA_MAX_WIN || !residual || (lam0 && cap < 0)) return VINA_E_ARG;
  if (!ctx->d_ba) return vn_fail(ctx, VINA_E_STATE, "no BA factors collected yet");
  if (win != ctx->cfg.win_size) return vn_fail(ctx, VINA_E_ARG, "win %d != LocalBA.win_size %d", win, ctx->cfg.win_size);
  int32_t n = 0;
  int r = vina_ba_count(ctx, &n);
  if (r) return r;
  const int nblk = ctx->sm_count * 2;
  const BaDone done = { ctx->d_ba_ticket, ctx->d_ba_flag, ++ctx->ba_seq };
  ctx->launches += launch_ba_residual(ctx->stream, ctx->d_ba, ctx->d_ba_n, reinterpret_cast<const PoseD*>(xs), win,
                                      ctx->sm_count, ctx->d_ba_map, ctx->d_ba_lam, done);  // block sums -> mapped memory
  r = vn_check_cuda(ctx, cudaGetLastError(), "k_ba_residual");
  if (r) return r;
  r = ba_wait(ctx, "k_ba_residual");
  if (r) return r;
  const double* part = ctx->h_ba_out;
  double s = 0.0;
  for (int b = 0; b < nblk; b++) s += part[b];  // block order: deterministic for a given factor order
  *residual = s;
  if (lam0 && n > 0)
  {
    CU(cudaMemcpyAsync(lam0, ctx->d_ba_lam, (size_t)(n < cap ? n : cap) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  return vn_check_cuda(ctx, cudaGetLastError(), "k_ba_residual");
}

extern "C" int vina_set_profiling(vina_ctx* ctx, int on)
{
  if (!ctx) return VINA_E_ARG;
  ctx->profiling = on != 0;
  return VINA_OK;
}

extern "C" int vina_get_timings(vina_ctx* ctx, vina_timings* t)
{
  if (!ctx || !t) return VINA_E_ARG;
  *t = ctx->tm;
  return VINA_OK;
}

// experiment hook: `reps` back-to-back launches of the accumulate kernel between two CUDA events
// (no host sync in between); variant bits switch parts of the kernel off (0 = product path).
extern "C" int vina_iekf_time_kernel(vina_ctx* ctx, const double R[9], const double p[3], int reps, int variant,
                                     int reset_cache, float* ms_per_launch)
{
  if (!ctx || !R || !p || reps < 1 || !ms_per_launch) return VINA_E_ARG;
  if (ctx->iekf_which < 0) return vn_fail(ctx, VINA_E_STATE, "vina_iekf_time_kernel before vina_iekf_begin");
  CU(cudaMemcpyAsync(ctx->d_iekf->R, R, 72, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_iekf->p, p, 24, cudaMemcpyHostToDevice, ctx->stream));
  IekfBatch bt;
  bt.mode = 0;  // sums stay in the device iterate
  bt.variant = variant & 0xff;
  const int blocks = (variant >> 8) > 0 ? (variant >> 8) : ctx->iekf_blocks;  // bits 8.. = grid override
  vn_iekf_fill_seq(ctx, &bt.s[0], false);
  cudaEventRecord(ctx->ev[10], ctx->stream);
  for (int r = 0; r < reps; r++)
  {
    if (reset_cache) launch_fill_int(ctx->stream, ctx->d_cache, -1, ctx->n_pv[ctx->iekf_which]);
    int e = launch_iekf(ctx->stream, bt, 1, blocks, false);
    if (e) return vn_check_cuda(ctx, (cudaError_t)e, "k_iekf launch");
  }
  cudaEventRecord(ctx->ev[11], ctx->stream);
  CU(cudaEventSynchronize(ctx->ev[11]));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev[10], ctx->ev[11]);
  *ms_per_launch = ms / reps;
  return VINA_OK;
}

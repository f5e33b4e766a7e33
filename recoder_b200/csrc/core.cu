// Error reporting and device queries shared by all entry points.
#include <stdarg.h>

#include "common.cuh"

#include <stdlib.h>
#include <string.h>

#include <thread>
#include <vector>

static thread_local char g_err[512] = "";
unsigned long long g_rcd_launches = 0;

void rcd_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int rcd_num_sms() {
  static int sms = 0;
  if (sms > 0) return sms;
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
  sms = v;
  return sms;
}

RCD_EXPORT int rcd_abi_version(void) { return RCD_ABI_VERSION; }
RCD_EXPORT const char* rcd_last_error(void) { return g_err; }
RCD_EXPORT int rcd_device_sms(void) {
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    rcd_set_error("rcd_device_sms: no CUDA device");
    return RCD_ERR_CUDA;
  }
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    rcd_set_error("rcd_device_sms: attribute query failed");
    return RCD_ERR_CUDA;
  }
  return v;
}

RCD_EXPORT long long rcd_launch_count(void) { return (long long)g_rcd_launches; }

// Host-side staging of one pool for the host-resident data path: copies the CSR rows `users` of a HOST matrix
// into (pinned) staging buffers in stored order — the work scipy's fancy row indexing does at
// recoder/data.py:66-81.  row_ptr_out int64[P+1] is the pool's own indptr.  Returns the pool's nnz (<0 on error).
RCD_EXPORT long long rcd_host_stage_rows(const int64_t* indptr_host, const int32_t* indices_host,
                                         const float* data_host, const int64_t* users_host, int pool_rows,
                                         long long num_users, long long capacity, int64_t* row_ptr_out_host,
                                         int32_t* indices_out_host, float* data_out_host) {
  if (!indptr_host || !indices_host || !data_host || !users_host || !row_ptr_out_host || !indices_out_host ||
      !data_out_host || pool_rows <= 0) {
    rcd_set_error("rcd_host_stage_rows: invalid argument");
    return RCD_ERR_INVALID;
  }
  // pass 1 (serial, P additions): validate and lay the rows out
  long long at = 0;
  row_ptr_out_host[0] = 0;
  for (int r = 0; r < pool_rows; ++r) {
    const int64_t u = users_host[r];
    if (u < 0 || u >= num_users) {
      rcd_set_error("rcd_host_stage_rows: user index %lld out of range", (long long)u);
      return RCD_ERR_INVALID;
    }
    at += (long long)(indptr_host[u + 1] - indptr_host[u]);
    if (at > capacity) {
      rcd_set_error("rcd_host_stage_rows: staging capacity %lld exceeded", capacity);
      return RCD_ERR_INVALID;
    }
    row_ptr_out_host[r + 1] = at;
  }
  // pass 2: the copies — one cache-missing row of the matrix per pool row (16 K rows per pool and rank in the 8-GPU
  // item-parallel mode, where this loop was what the step waited for) — on a few threads (RCD_STAGE_THREADS, default 4)
  auto copy_range = [&](int r0, int r1) {
    for (int r = r0; r < r1; ++r) {
      const int64_t s0 = indptr_host[users_host[r]];
      const long long o = row_ptr_out_host[r], len = row_ptr_out_host[r + 1] - o;
      memcpy(indices_out_host + o, indices_host + s0, (size_t)len * sizeof(int32_t));
      memcpy(data_out_host + o, data_host + s0, (size_t)len * sizeof(float));
    }
  };
  static int n_threads = 0;
  if (n_threads == 0) {
    const char* e = getenv("RCD_STAGE_THREADS");
    const int v = e ? atoi(e) : 4;
    n_threads = (v >= 1 && v <= 64) ? v : 4;
  }
  const int nt = (pool_rows >= 2048 && n_threads > 1) ? n_threads : 1;
  if (nt == 1) {
    copy_range(0, pool_rows);
  } else {
    std::vector<std::thread> th;
    const int per = (pool_rows + nt - 1) / nt;
    for (int t = 1; t < nt; ++t) {
      const int r0 = t * per, r1 = (t + 1) * per < pool_rows ? (t + 1) * per : pool_rows;
      if (r0 < r1) th.emplace_back(copy_range, r0, r1);
    }
    copy_range(0, per < pool_rows ? per : pool_rows);
    for (auto& x : th) x.join();
  }
  return at;
}

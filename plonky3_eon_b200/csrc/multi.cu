// Multi-device context: ONE process drives several B200s behind the same C ABI (include/eon_kzg.h, eon_mctx_*).
//
// The reference prover is a single process that calls pcs.commit / get_evaluations_on_domain / open on whole
// matrices (eon-uni-stark/src/prover.rs:186-187,307-322,371-372,424-442) and KzgPcs walks the columns of a matrix
// one after the other (kzg/src/pcs.rs:244-249,311-318).  Every column's iDFT, LDE, MSM and opening is independent,
// so the matrix shards by COLUMNS: device g owns the contiguous column range column_shard(width, G, g), reads it
// straight out of the caller's row-major host matrix with strided copies (row pitch = full width) and writes its
// slice of every result straight back; no collective.  A single big MSM (G1::multi_exp, bn254/src/curve.rs:158-180)
// with fewer columns than devices shards by POINT INDEX instead: device g sums SRS[first_g ..) with window tables
// built for exactly that range, the G partial sums travel to device 0 as peer copies over NVLink and one launch
// adds them (EC addition is not a collective reduction op).
//
// One host thread per device issues that device's work (the single-device entry points synchronise their own
// stream before they return, so devices only overlap when driven from different threads).
#include <condition_variable>
#include <functional>
#include <thread>

#include "common.cuh"

using namespace eon;

namespace {

// persistent workers: run(n, f) executes f(0..n-1) concurrently (f(0) on the calling thread)
class Pool {
 public:
  explicit Pool(int workers) : jobs_(workers), gen_(workers, 0), stop_(false) {
    for (int i = 0; i < workers; i++) threads_.emplace_back([this, i] { loop(i); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> l(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  void run(int n, const std::function<void(int)>& f) {
    int started = 0;
    {
      std::lock_guard<std::mutex> l(mu_);
      for (int i = 1; i < n; i++) {
        jobs_[i - 1] = [&f, i] { f(i); };
        gen_[i - 1]++;
        started++;
      }
      pending_ = started;
    }
    cv_.notify_all();
    if (n > 0) f(0);
    std::unique_lock<std::mutex> l(mu_);
    done_.wait(l, [this] { return pending_ == 0; });
  }

 private:
  void loop(int i) {
    unsigned long seen = 0;
    for (;;) {
      std::function<void()> job;
      {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [&] { return stop_ || gen_[i] != seen; });
        if (stop_) return;
        seen = gen_[i];
        job = jobs_[i];
      }
      job();
      {
        std::lock_guard<std::mutex> l(mu_);
        pending_--;
      }
      done_.notify_all();
    }
  }
  std::vector<std::function<void()>> jobs_;
  std::vector<unsigned long> gen_;
  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  int pending_ = 0;
  bool stop_;
};

struct MShard {
  int dev;  // index into eon_mctx::ctx
  eon_handle h;
  size_t c0, w;  // column range [c0, c0 + w) of the matrix
};
struct MHandle {
  std::vector<MShard> shards;
  size_t rows, width;
  unsigned log_h;
};

}  // namespace

struct eon_mctx {
  std::vector<eon_ctx*> ctx;
  std::vector<cudaStream_t> streams;  // owned: one non-blocking stream per shard context
  std::vector<cudaEvent_t> ev;
  std::mutex mu;
  std::string last_error;
  std::map<eon_handle, MHandle> handles;
  eon_handle next_handle = 1;
  Pool* pool = nullptr;
  G1Affine* d_gather = nullptr;  // on ctx[0]'s device: the shards' partial sums, [shard][column]
  size_t gather_cap = 0;
};

namespace {

int mfail(eon_mctx* m, int code, const std::string& msg) {
  if (m) m->last_error = msg;
  return code;
}

// balanced contiguous ranges: the first (total % parts) parts get one more
void shard_range(size_t total, size_t parts, size_t i, size_t* first, size_t* count) {
  const size_t base = total / parts, extra = total % parts;
  *first = i * base + std::min(i, extra);
  *count = base + (i < extra ? 1 : 0);
}

// f(i) -> status for every shard context in parallel; first failure wins, its message is kept
int par(eon_mctx* m, int n, const std::function<int(int)>& f) {
  std::vector<int> rc((size_t)std::max(n, 1), EON_OK);
  m->pool->run(n, [&](int i) { rc[(size_t)i] = f(i); });
  for (int i = 0; i < n; i++)
    if (rc[(size_t)i] != EON_OK) {
      char b[64];
      snprintf(b, sizeof(b), "device %d: ", m->ctx[(size_t)i]->device);
      m->last_error = std::string(b) + eon_last_error(m->ctx[(size_t)i]);
      return rc[(size_t)i];
    }
  return EON_OK;
}

int find(eon_mctx* m, eon_handle h, MHandle** out) {
  auto it = m->handles.find(h);
  if (it == m->handles.end()) return mfail(m, EON_ERR_BAD_HANDLE, "unknown prover-data handle");
  *out = &it->second;
  return EON_OK;
}

eon_handle reg(eon_mctx* m, const MHandle& mh) {
  eon_handle id = m->next_handle++;
  m->handles[id] = mh;
  return id;
}

// the devices that get columns of a `width`-column matrix (all of them once width >= G; a zero-width matrix still
// goes to device 0 so that the shape / SRS checks of the single-device entry point apply)
int active_shards(const eon_mctx* m, size_t width) { return (int)std::max<size_t>(1, std::min(width, m->ctx.size())); }

}  // namespace

extern "C" {

void eon_mctx_destroy(eon_mctx* m);

int eon_mctx_create(const int* devices, int n, eon_mctx** out) {
  if (!out) return EON_ERR_BAD_ARG;
  *out = nullptr;
  if (!devices || n < 1 || n > 64) return EON_ERR_BAD_ARG;
  eon_mctx* m = new eon_mctx();
  for (int i = 0; i < n; i++) {
    cudaStream_t st = nullptr;
    eon_ctx* c = nullptr;
    int rc = EON_ERR_CUDA;
    if (cudaSetDevice(devices[i]) == cudaSuccess && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess)
      rc = eon_ctx_create(devices[i], st, &c);
    if (rc != EON_OK) {
      if (st) cudaStreamDestroy(st);
      eon_mctx_destroy(m);
      return rc;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    m->ctx.push_back(c);
    m->streams.push_back(st);
    m->ev.push_back(e);
  }
  // NVLink peer access towards device 0 (the partial sums of an index-range MSM are pushed there); a pair that
  // cannot be mapped falls back to staging through the host inside cudaMemcpyPeerAsync
  for (int i = 1; i < n; i++) {
    if (devices[i] == devices[0]) continue;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, devices[i], devices[0]) == cudaSuccess && can) {
      cudaSetDevice(devices[i]);
      cudaError_t e = cudaDeviceEnablePeerAccess(devices[0], 0);
      if (e != cudaSuccess) cudaGetLastError();  // already enabled by someone else in this process
    }
  }
  m->pool = new Pool(n - 1);
  *out = m;
  return EON_OK;
}

void eon_mctx_destroy(eon_mctx* m) {
  if (!m) return;
  delete m->pool;
  if (m->d_gather && !m->ctx.empty()) {
    cudaSetDevice(m->ctx[0]->device);
    cudaFree(m->d_gather);
  }
  for (size_t i = 0; i < m->ctx.size(); i++) {
    const int dev = m->ctx[i]->device;
    eon_ctx_destroy(m->ctx[i]);  // frees every prover-data handle of that device as well
    cudaSetDevice(dev);
    if (m->ev[i]) cudaEventDestroy(m->ev[i]);
    cudaStreamDestroy(m->streams[i]);
  }
  delete m;
}

const char* eon_mctx_last_error(const eon_mctx* m) { return m ? m->last_error.c_str() : "null context"; }
int eon_mctx_device_count(const eon_mctx* m) { return m ? (int)m->ctx.size() : 0; }
eon_ctx* eon_mctx_ctx(eon_mctx* m, int i) { return (m && i >= 0 && (size_t)i < m->ctx.size()) ? m->ctx[(size_t)i] : nullptr; }

uint64_t eon_mctx_launch_count(const eon_mctx* m) {
  uint64_t t = 0;
  if (m)
    for (eon_ctx* c : m->ctx) t += eon_ctx_launch_count(c);
  return t;
}

// ---- SRS: replicated on every device (2^24 points = 1 GiB; kzg/src/params.rs:57-77) -----------------------------
int eon_mctx_srs_generate_unsafe(eon_mctx* m, const uint64_t alpha[4], size_t n) {
  if (!m) return EON_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(m->mu);
  return par(m, (int)m->ctx.size(), [&](int i) { return eon_srs_generate_unsafe(m->ctx[(size_t)i], alpha, n); });
}
int eon_mctx_srs_load_affine(eon_mctx* m, const uint64_t* h_xy, size_t n) {
  if (!m) return EON_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(m->mu);
  return par(m, (int)m->ctx.size(), [&](int i) { return eon_srs_load_affine(m->ctx[(size_t)i], h_xy, n); });
}
int eon_mctx_srs_load_compressed(eon_mctx* m, const uint8_t* h_in, size_t n, int enc, size_t* bad_index) {
  if (!m) return EON_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(m->mu);
  std::vector<size_t> bad(m->ctx.size(), (size_t)-1);
  int rc = par(m, (int)m->ctx.size(),
               [&](int i) { return eon_srs_load_compressed(m->ctx[(size_t)i], h_in, n, enc, &bad[(size_t)i]); });
  if (bad_index) *bad_index = bad[0];
  return rc;
}
size_t eon_mctx_srs_size(const eon_mctx* m) { return m ? eon_srs_size(m->ctx[0]) : 0; }
int eon_mctx_srs_read(eon_mctx* m, size_t first, size_t n, uint64_t* h_xy) {
  if (!m) return EON_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(m->mu);
  return par(m, 1, [&](int) { return eon_srs_read(m->ctx[0], first, n, h_xy); });
}

// ---- TwoAdicSubgroupDft<Fr> on whole host matrices, columns sharded ---------------------------------------------
static int mctx_dft(eon_mctx* m, int kind, const uint64_t* h_in, uint64_t* h_out, unsigned log_h, size_t width,
                    unsigned added_bits, const uint64_t shift[4]) {
  if (!m) return EON_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(m->mu);
  const int k = active_shards(m, width);
  return par(m, k, [&](int i) {
    size_t c0, w;
    shard_range(width, (size_t)k, (size_t)i, &c0, &w);
    return dft_host(m->ctx[(size_t)i], kind, h_in ? h_in + c0 * 4 : nullptr, width, h_out ? h_out + c0 * 4 : nullptr, width,
                    log_h, w, added_bits, shift);
  });
}
int eon_mctx_dft_batch(eon_mctx* m, const uint64_t* h_in, uint64_t* h_out, unsigned log_h, size_t width) {
  return mctx_dft(m, DFT_PLAIN, h_in, h_out, log_h, width, 0, nullptr);
}
int eon_mctx_coset_dft_batch(eon_mctx* m, const uint64_t* h_in, uint64_t* h_out, unsigned log_h, size_t width,
                             const uint64_t shift[4]) {
  return mctx_dft(m, DFT_COSET, h_in, h_out, log_h, width, 0, shift);
}
int eon_mctx_idft_batch(eon_mctx* m, const uint64_t* h_in, uint64_t* h_out, unsigned log_h, size_t width) {
  return mctx_dft(m, DFT_INV, h_in, h_out, log_h, width, 0, nullptr);
}
int eon_mctx_coset_idft_batch(eon_mctx* m, const uint64_t* h_in, uint64_t* h_out, unsigned log_h, size_t width,
                              const uint64_t shift[4]) {
  return mctx_dft(m, DFT_COSET_INV, h_in, h_out, log_h, width, 0, shift);
}
int eon_mctx_coset_lde_batch(eon_mctx* m, const uint64_t* h_in, uint64_t* h_out, unsigned log_h, size_t width,
                             unsigned added_bits, const uint64_t shift[4]) {
  return mctx_dft(m, DFT_COSET_LDE, h_in, h_out, log_h, width, added_bits, shift);
}

// ---- KzgPcs ---------------------------------------------------------------------------------------------------------
static int mctx_commit(eon_mctx* m, const uint64_t* h_evals, unsigned log_h, size_t width, const uint64_t shift[4],
                       uint64_t* h_commit_xy, eon_handle* out_handle, unsigned lde_log_size, const uint64_t* lde_shift,
                       uint64_t* h_lde_out) {
  if (!m || !out_handle) return EON_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(m->mu);
  *out_handle = 0;
  const int k = active_shards(m, width);
  MHandle mh;
  mh.rows = (size_t)1 << (log_h > 28 ? 28 : log_h);
  mh.log_h = log_h;
  mh.width = width;
  mh.shards.resize((size_t)k);
  for (int i = 0; i < k; i++) {
    mh.shards[(size_t)i].dev = i;
    mh.shards[(size_t)i].h = 0;
    shard_range(width, (size_t)k, (size_t)i, &mh.shards[(size_t)i].c0, &mh.shards[(size_t)i].w);
  }
  int rc = par(m, k, [&](int i) {
    MShard& s = mh.shards[(size_t)i];
    const uint64_t* in = h_evals ? h_evals + s.c0 * 4 : nullptr;
    uint64_t* cm = h_commit_xy ? h_commit_xy + s.c0 * 8 : nullptr;
    if (lde_log_size)
      return eon_kzg_commit_lde_ld(m->ctx[(size_t)i], in, width, log_h, s.w, shift, cm, &s.h, lde_log_size, lde_shift,
                                   h_lde_out ? h_lde_out + s.c0 * 4 : nullptr, width);
    return eon_kzg_commit_ld(m->ctx[(size_t)i], in, width, log_h, s.w, shift, cm, &s.h);
  });
  if (rc != EON_OK) {
    for (auto& s : mh.shards)
      if (s.h) eon_handle_free(m->ctx[(size_t)s.dev], s.h);
    return rc;
  }
  *out_handle = reg(m, mh);
  return EON_OK;
}

int eon_mctx_kzg_commit(eon_mctx* m, const uint64_t* h_evals, unsigned log_h, size_t width, const uint64_t shift[4],
                        uint64_t* h_commit_xy, eon_handle* out_handle) {
  return mctx_commit(m, h_evals, log_h, width, shift, h_commit_xy, out_handle, 0, nullptr, nullptr);
}
int eon_mctx_kzg_commit_lde(eon_mctx* m, const uint64_t* h_evals, unsigned log_h, size_t width, const uint64_t shift[4],
                            uint64_t* h_commit_xy, eon_handle* out_handle, unsigned lde_log_size,
                            const uint64_t lde_shift[4], uint64_t* h_lde_out) {
  if (lde_log_size == 0) return EON_ERR_BAD_ARG;
  return mctx_commit(m, h_evals, log_h, width, shift, h_commit_xy, out_handle, lde_log_size, lde_shift, h_lde_out);
}

int eon_mctx_kzg_commit_coeffs(eon_mctx* m, const uint64_t* h_coeffs, size_t rows, size_t width, uint64_t* h_commit_xy,
                               eon_handle* out_handle) {
  if (!m || !out_handle) return EON_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(m->mu);
  *out_handle = 0;
  const int k = active_shards(m, width);
  MHandle mh;
  mh.rows = rows;
  mh.width = width;
  mh.log_h = NOT_POW2;
  if (rows && (rows & (rows - 1)) == 0) {
    mh.log_h = 0;
    while (((size_t)1 << mh.log_h) < rows) mh.log_h++;
  }
  mh.shards.resize((size_t)k);
  for (int i = 0; i < k; i++) {
    mh.shards[(size_t)i].dev = i;
    mh.shards[(size_t)i].h = 0;
    shard_range(width, (size_t)k, (size_t)i, &mh.shards[(size_t)i].c0, &mh.shards[(size_t)i].w);
  }
  int rc = par(m, k, [&](int i) {
    MShard& s = mh.shards[(size_t)i];
    return kzg_commit_coeffs_host_ld(m->ctx[(size_t)i], h_coeffs ? h_coeffs + s.c0 * 4 : nullptr, width, rows, s.w,
                                     h_commit_xy ? h_commit_xy + s.c0 * 8 : nullptr, &s.h);
  });
  if (rc != EON_OK) {
    for (auto& s : mh.shards)
      if (s.h) eon_handle_free(m->ctx[(size_t)s.dev], s.h);
    return rc;
  }
  *out_handle = reg(m, mh);
  return EON_OK;
}

// Pcs::commit_quotient (commit/src/pcs.rs:82-102): the 2^log_chunks chunks go to the devices as contiguous ranges
// (a chunk is a pitched view of the host matrix: its rows r = i mod 2^log_chunks are one strided copy)
int eon_mctx_kzg_commit_quotient(eon_mctx* m, const uint64_t* h_evals, unsigned log_size, size_t width,
                                 unsigned log_chunks, const uint64_t shift[4], uint64_t* h_commit_xy,
                                 eon_handle* out_handles) {
  if (!m || !out_handles) return EON_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(m->mu);
  if (log_chunks > 16) return mfail(m, EON_ERR_BAD_ARG, "too many quotient chunks");
  const size_t nchunks = (size_t)1 << log_chunks;
  for (size_t i = 0; i < nchunks; i++) out_handles[i] = 0;
  const int k = (int)std::min(nchunks, m->ctx.size());
  std::vector<eon_handle> hs(nchunks, 0);
  int rc = par(m, k, [&](int i) {
    size_t k0, nk;
    shard_range(nchunks, (size_t)k, (size_t)i, &k0, &nk);
    return kzg_commit_quotient_range_host(m->ctx[(size_t)i], h_evals, log_size, width, log_chunks, k0, nk, shift,
                                          h_commit_xy ? h_commit_xy + k0 * width * 8 : nullptr, hs.data() + k0);
  });
  if (rc != EON_OK) {
    for (int i = 0; i < k; i++) {
      size_t k0, nk;
      shard_range(nchunks, (size_t)k, (size_t)i, &k0, &nk);
      for (size_t j = k0; j < k0 + nk; j++)
        if (hs[j]) eon_handle_free(m->ctx[(size_t)i], hs[j]);
    }
    return rc;
  }
  for (int i = 0; i < k; i++) {
    size_t k0, nk;
    shard_range(nchunks, (size_t)k, (size_t)i, &k0, &nk);
    for (size_t j = k0; j < k0 + nk; j++) {
      MHandle mh;
      mh.log_h = log_size - log_chunks;
      mh.rows = (size_t)1 << mh.log_h;
      mh.width = width;
      mh.shards.push_back(MShard{i, hs[j], 0, width});
      out_handles[j] = reg(m, mh);
    }
  }
  return EON_OK;
}

int eon_mctx_handle_dims(eon_mctx* m, eon_handle h, unsigned* log_h, size_t* width) {
  if (!m) return EON_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(m->mu);
  MHandle* mh;
  EON_TRY(find(m, h, &mh));
  if (log_h) *log_h = mh->log_h;
  if (width) *width = mh->width;
  return EON_OK;
}

int eon_mctx_handle_free(eon_mctx* m, eon_handle h) {
  if (!m) return EON_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(m->mu);
  MHandle* mh;
  EON_TRY(find(m, h, &mh));
  int rc = EON_OK;
  for (auto& s : mh->shards) {
    int r = eon_handle_free(m->ctx[(size_t)s.dev], s.h);
    if (r != EON_OK && rc == EON_OK) {
      rc = r;
      m->last_error = eon_last_error(m->ctx[(size_t)s.dev]);
    }
  }
  m->handles.erase(h);
  return rc;
}

int eon_mctx_kzg_read_coeffs(eon_mctx* m, eon_handle h, uint64_t* h_out) {
  if (!m) return EON_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(m->mu);
  MHandle* mh;
  EON_TRY(find(m, h, &mh));
  if (mh->rows * mh->width == 0) return EON_OK;
  if (!h_out) return mfail(m, EON_ERR_BAD_ARG, "null output");
  const MHandle& H = *mh;
  return par(m, (int)H.shards.size(), [&](int i) {
    const MShard& s = H.shards[(size_t)i];
    if (s.w == 0) return (int)EON_OK;
    if (s.w == H.width) return eon_kzg_read_coeffs(m->ctx[(size_t)s.dev], s.h, h_out);
    std::vector<uint64_t> tmp(H.rows * s.w * 4);
    int rc = eon_kzg_read_coeffs(m->ctx[(size_t)s.dev], s.h, tmp.data());
    if (rc != EON_OK) return rc;
    for (size_t r = 0; r < H.rows; r++)
      memcpy(h_out + (r * H.width + s.c0) * 4, tmp.data() + r * s.w * 4, s.w * 32);
    return (int)EON_OK;
  });
}

int eon_mctx_kzg_evals_on_coset(eon_mctx* m, eon_handle h, unsigned log_size, const uint64_t shift[4], uint64_t* h_out) {
  if (!m) return EON_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(m->mu);
  MHandle* mh;
  EON_TRY(find(m, h, &mh));
  const MHandle& H = *mh;
  return par(m, (int)H.shards.size(), [&](int i) {
    const MShard& s = H.shards[(size_t)i];
    return eon_kzg_evals_on_coset_ld(m->ctx[(size_t)s.dev], s.h, log_size, shift, h_out ? h_out + s.c0 * 4 : nullptr,
                                     H.width);
  });
}

// open (kzg/src/pcs.rs:289-335): every device opens its column shard of every matrix with ONE batched call; the
// per-device results are scattered into the [matrix][point][column] layout of the single-device entry point
int eon_mctx_kzg_open_batch(eon_mctx* m, size_t nmat, const eon_handle* handles, const size_t* npoints,
                            const uint64_t* h_points, uint64_t* h_values, uint64_t* h_witness_xy) {
  if (!m) return EON_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(m->mu);
  if (nmat == 0) return EON_OK;
  if (!handles || !npoints) return mfail(m, EON_ERR_BAD_ARG, "null handle / point-count array");
  std::vector<const MHandle*> mhs(nmat);
  std::vector<size_t> pt0(nmat), out0(nmat);  // first point / first output slot of every matrix
  size_t total_points = 0, total_out = 0;
  for (size_t i = 0; i < nmat; i++) {
    MHandle* mh;
    EON_TRY(find(m, handles[i], &mh));
    mhs[i] = mh;
    pt0[i] = total_points;
    out0[i] = total_out;
    total_points += npoints[i];
    total_out += npoints[i] * mh->width;
  }
  if (total_out == 0) return EON_OK;
  if (!h_points || !h_values || !h_witness_xy) return mfail(m, EON_ERR_BAD_ARG, "null buffer");
  const int G = (int)m->ctx.size();
  struct Job {
    std::vector<eon_handle> hs;
    std::vector<size_t> np, mat, c0, w;
    std::vector<uint64_t> pts, vals, wits;
  };
  std::vector<Job> jobs((size_t)G);
  for (size_t i = 0; i < nmat; i++)
    for (const MShard& s : mhs[i]->shards) {
      if (s.w == 0 || npoints[i] == 0) continue;
      Job& j = jobs[(size_t)s.dev];
      j.hs.push_back(s.h);
      j.np.push_back(npoints[i]);
      j.mat.push_back(i);
      j.c0.push_back(s.c0);
      j.w.push_back(s.w);
      j.pts.insert(j.pts.end(), h_points + pt0[i] * 4, h_points + (pt0[i] + npoints[i]) * 4);
    }
  int rc = par(m, G, [&](int d) {
    Job& j = jobs[(size_t)d];
    if (j.hs.empty()) return (int)EON_OK;
    size_t tot = 0;
    for (size_t q = 0; q < j.hs.size(); q++) tot += j.np[q] * j.w[q];
    j.vals.resize(tot * 4);
    j.wits.resize(tot * 8);
    int r = eon_kzg_open_batch(m->ctx[(size_t)d], j.hs.size(), j.hs.data(), j.np.data(), j.pts.data(), j.vals.data(),
                               j.wits.data());
    if (r != EON_OK) return r;
    size_t k = 0;
    for (size_t q = 0; q < j.hs.size(); q++) {
      const size_t W = mhs[j.mat[q]]->width;
      for (size_t p = 0; p < j.np[q]; p++) {
        const size_t dst = out0[j.mat[q]] + p * W + j.c0[q];
        memcpy(h_values + dst * 4, j.vals.data() + k * 4, j.w[q] * 32);
        memcpy(h_witness_xy + dst * 8, j.wits.data() + k * 8, j.w[q] * 64);
        k += j.w[q];
      }
    }
    return (int)EON_OK;
  });
  return rc;
}

// ---- G1::multi_exp over the resident SRS ------------------------------------------------------------------------------
// ncols >= devices: columns sharded (as a commit).  Fewer columns than devices: the POINTS are sharded by index
// range; every device keeps window tables for exactly its range (built on first use, sized by the cost model for
// the shard, not for the whole SRS), leaves its ncols partial sums in device memory and pushes them to device 0
// as a peer copy; one launch there adds the shards' sums of all columns.
int eon_mctx_msm_srs(eon_mctx* m, const uint64_t* h_scalars, size_t n, size_t ncols, size_t ld, uint64_t* h_out_xy) {
  if (!m) return EON_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(m->mu);
  if (ncols == 0) return EON_OK;
  if (ld < ncols) return mfail(m, EON_ERR_BAD_ARG, "ld < ncols");
  if (!h_out_xy || (n && !h_scalars)) return mfail(m, EON_ERR_BAD_ARG, "null buffer");
  const size_t G = m->ctx.size();
  if (ncols >= G || n < ((size_t)1 << 15)) {
    const int k = active_shards(m, ncols);
    return par(m, k, [&](int i) {
      size_t c0, w;
      shard_range(ncols, (size_t)k, (size_t)i, &c0, &w);
      return msm_srs_host_ld(m->ctx[(size_t)i], h_scalars + c0 * 4, n, w, ld, h_out_xy + c0 * 8);
    });
  }
  if (n > eon_srs_size(m->ctx[0])) return mfail(m, EON_ERR_SRS_TOO_SHORT, "DegreeTooLarge: polynomial longer than the SRS");
  const int k = (int)G;
  eon_ctx* c0 = m->ctx[0];
  if (m->gather_cap < (size_t)k * ncols) {
    if (cudaSetDevice(c0->device) != cudaSuccess) return mfail(m, EON_ERR_CUDA, "cudaSetDevice failed");
    if (m->d_gather) {
      cudaStreamSynchronize(c0->stream);
      cudaFree(m->d_gather);
      m->d_gather = nullptr;
    }
    if (cudaMalloc(&m->d_gather, (size_t)k * ncols * sizeof(G1Affine)) != cudaSuccess)
      return mfail(m, EON_ERR_OOM, "gather buffer allocation failed");
    m->gather_cap = (size_t)k * ncols;
  }
  int rc = par(m, k, [&](int i) {
    eon_ctx* c = m->ctx[(size_t)i];
    size_t first, cnt;
    shard_range(n, (size_t)k, (size_t)i, &first, &cnt);
    if (cnt >= ((size_t)1 << 14) && (c->rng_first != first || c->rng_n != cnt)) {
      int r = eon_srs_set_range_tables(c, first, cnt, 0);
      if (r != EON_OK) return r;
    }
    const G1Affine* part = nullptr;
    int r = msm_srs_range_host_partial(c, h_scalars + first * ld * 4, first, cnt, ncols, ld, &part);
    if (r != EON_OK) return r;
    if (cudaSetDevice(c->device) != cudaSuccess) return fail(c, EON_ERR_CUDA, "cudaSetDevice failed");
    cudaError_t e = cudaMemcpyPeerAsync(m->d_gather + (size_t)i * ncols, c0->device, part, c->device,
                                        ncols * sizeof(G1Affine), c->stream);
    if (e == cudaSuccess) e = cudaEventRecord(m->ev[(size_t)i], c->stream);
    if (e != cudaSuccess) return fail(c, EON_ERR_CUDA, std::string("partial-sum push failed: ") + cudaGetErrorString(e));
    return (int)EON_OK;
  });
  if (rc != EON_OK) {
    for (eon_ctx* c : m->ctx) eon_ctx_sync(c);
    return rc;
  }
  if (cudaSetDevice(c0->device) != cudaSuccess) return mfail(m, EON_ERR_CUDA, "cudaSetDevice failed");
  for (int i = 0; i < k; i++)
    if (cudaStreamWaitEvent(c0->stream, m->ev[(size_t)i], 0) != cudaSuccess) return mfail(m, EON_ERR_CUDA, "event wait failed");
  rc = eon_g1_sum_cols_dev(c0, (const uint64_t*)m->d_gather, (size_t)k, ncols, h_out_xy);
  if (rc != EON_OK) m->last_error = eon_last_error(c0);
  return rc;
}

// explicit bases (no tables): the points are sharded by index range, partial sums added on device 0
int eon_mctx_msm_points(eon_mctx* m, const uint64_t* h_points_xy, const uint64_t* h_scalars, size_t n, uint64_t* h_out_xy) {
  if (!m) return EON_ERR_BAD_ARG;
  std::lock_guard<std::mutex> lk(m->mu);
  if (!h_out_xy || (n && (!h_points_xy || !h_scalars))) return mfail(m, EON_ERR_BAD_ARG, "null buffer");
  const size_t G = m->ctx.size();
  const int k = (int)std::max<size_t>(1, std::min(G, n >> 12));
  std::vector<uint64_t> parts((size_t)k * 8);
  int rc = par(m, k, [&](int i) {
    size_t first, cnt;
    shard_range(n, (size_t)k, (size_t)i, &first, &cnt);
    return eon_msm_points(m->ctx[(size_t)i], h_points_xy + first * 8, h_scalars + first * 4, cnt, parts.data() + (size_t)i * 8);
  });
  if (rc != EON_OK) return rc;
  if (k == 1) {
    memcpy(h_out_xy, parts.data(), 64);
    return EON_OK;
  }
  rc = eon_g1_sum(m->ctx[0], parts.data(), (size_t)k, h_out_xy);
  if (rc != EON_OK) m->last_error = eon_last_error(m->ctx[0]);
  return rc;
}

}  // extern "C"

// dnbc_comm.cu -- the one exchange step of the path, behind the C ABI.
//
// The reference learns in parallel over model PARAMETERS: three Spark parallelize -> map -> collect
// rounds through the driver (DynamicNaiveBayesianClassifier.scala:231-232, 241-243, 250-252).  Here
// sequences shard over GPUs and a single ncclAllReduce(sum, ncclDouble) of the packed statistics
// buffer per EM iteration merges the shards over NVLink / NVSwitch; it is enqueued on the context's
// own stream right behind the statistics kernels, so nothing synchronises with the host between the
// E-step and the M-step.  The buffer is 312 KB at the scale-out shape: latency-bound.
#include <nccl.h>
#include <string.h>

#include <vector>

#include "dnbc_internal.cuh"

#define NCCL_TRY(c, expr)                                                                         \
  do {                                                                                            \
    ncclResult_t r__ = (expr);                                                                    \
    if (r__ != ncclSuccess) return dnbc_fail((c), DNBC_ECOMM, "%s: %s", #expr, ncclGetErrorString(r__)); \
  } while (0)

static_assert(DNBC_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "unique id size");

extern "C" int dnbc_comm_unique_id(uint8_t id[DNBC_COMM_ID_BYTES]) {
  if (!id) return DNBC_EINVAL;
  ncclUniqueId u;
  ncclResult_t r = ncclGetUniqueId(&u);
  if (r != ncclSuccess) return dnbc_fail(nullptr, DNBC_ECOMM, "ncclGetUniqueId: %s", ncclGetErrorString(r));
  memcpy(id, u.internal, DNBC_COMM_ID_BYTES);
  return DNBC_OK;
}

extern "C" int dnbc_comm_init_rank(dnbc_ctx *c, int n_ranks, int rank, const uint8_t id[DNBC_COMM_ID_BYTES]) {
  if (!c || !id) return DNBC_EINVAL;
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return dnbc_fail(c, DNBC_EINVAL, "comm_init_rank: bad rank %d of %d", rank, n_ranks);
  if (c->comm) return dnbc_fail(c, DNBC_ESTATE, "comm_init_rank: the context already has a communicator");
  CU_TRY(c, cudaSetDevice(c->device));
  ncclUniqueId u;
  memcpy(u.internal, id, DNBC_COMM_ID_BYTES);
  ncclComm_t comm = nullptr;
  NCCL_TRY(c, ncclCommInitRank(&comm, n_ranks, u, rank));
  c->comm = comm;
  c->comm_ranks = n_ranks;
  c->comm_rank = rank;
  return DNBC_OK;
}

extern "C" int dnbc_comm_init_all(dnbc_ctx **ctxs, int n) {
  if (!ctxs || n < 1) return DNBC_EINVAL;
  std::vector<int> devs(n);
  for (int i = 0; i < n; ++i) {
    if (!ctxs[i]) return DNBC_EINVAL;
    if (ctxs[i]->comm) return dnbc_fail(ctxs[i], DNBC_ESTATE, "comm_init_all: context %d already has a communicator", i);
    if (ctxs[i]->stats_size != ctxs[0]->stats_size) return dnbc_fail(ctxs[0], DNBC_EINVAL, "comm_init_all: contexts differ in model shape");
    devs[i] = ctxs[i]->device;
    for (int k = 0; k < i; ++k)
      if (devs[k] == devs[i]) return dnbc_fail(ctxs[0], DNBC_EINVAL, "comm_init_all: device %d appears twice", devs[i]);
  }
  std::vector<ncclComm_t> comms(n, nullptr);
  NCCL_TRY(ctxs[0], ncclCommInitAll(comms.data(), n, devs.data()));
  for (int i = 0; i < n; ++i) {
    ctxs[i]->comm = comms[i];
    ctxs[i]->comm_ranks = n;
    ctxs[i]->comm_rank = i;
  }
  return DNBC_OK;
}

extern "C" int dnbc_comm_info(const dnbc_ctx *c, int *n_ranks, int *rank) {
  if (!c) return DNBC_EINVAL;
  if (n_ranks) *n_ranks = c->comm ? c->comm_ranks : 1;
  if (rank) *rank = c->comm ? c->comm_rank : 0;
  return DNBC_OK;
}

extern "C" int dnbc_comm_destroy(dnbc_ctx *c) {
  if (!c) return DNBC_EINVAL;
  if (c->comm) {
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    ncclCommDestroy((ncclComm_t)c->comm);
    c->comm = nullptr;
    c->comm_ranks = 1;
    c->comm_rank = 0;
  }
  return DNBC_OK;
}

extern "C" int dnbc_em_allreduce(dnbc_ctx *c) {
  if (!c) return DNBC_EINVAL;
  if (!c->comm || c->comm_ranks == 1) return DNBC_OK;
  CU_TRY(c, cudaSetDevice(c->device));
  NCCL_TRY(c, ncclAllReduce(c->d_stats, c->d_stats, (size_t)c->stats_size, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream));
  c->collectives++;
  return DNBC_OK;
}

extern "C" int dnbc_em_allreduce_group(dnbc_ctx **ctxs, int n) {
  if (!ctxs || n < 1) return DNBC_EINVAL;
  if (n == 1 && (!ctxs[0]->comm || ctxs[0]->comm_ranks == 1)) return DNBC_OK;
  for (int i = 0; i < n; ++i)
    if (!ctxs[i] || !ctxs[i]->comm || ctxs[i]->comm_ranks != n)
      return dnbc_fail(ctxs[0], DNBC_ESTATE, "allreduce_group: contexts were not joined by dnbc_comm_init_all");
  NCCL_TRY(ctxs[0], ncclGroupStart());
  for (int i = 0; i < n; ++i) {
    dnbc_ctx *c = ctxs[i];
    ncclResult_t r = ncclAllReduce(c->d_stats, c->d_stats, (size_t)c->stats_size, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream);
    if (r != ncclSuccess) {
      ncclGroupEnd();
      return dnbc_fail(ctxs[0], DNBC_ECOMM, "ncclAllReduce (context %d): %s", i, ncclGetErrorString(r));
    }
    c->collectives++;
  }
  NCCL_TRY(ctxs[0], ncclGroupEnd());
  return DNBC_OK;
}

extern "C" int dnbc_em_iterate_sharded(dnbc_ctx *c, int n_iters, int64_t total_sequences, double *loglik_out) {
  if (!c) return DNBC_EINVAL;
  if (total_sequences <= 0) return dnbc_fail(c, DNBC_EINVAL, "em_iterate_sharded: total_sequences must be positive");
  for (int i = 0; i < n_iters; ++i) {
    int rc = dnbc_em_estep(c, 0);       // an empty shard contributes zeros
    if (rc) return rc;
    rc = dnbc_em_allreduce(c);
    if (rc) return rc;
    if (loglik_out) CU_TRY(c, cudaMemcpyAsync(loglik_out + i, c->d_stats, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    rc = dnbc_em_mstep(c, total_sequences);
    if (rc) return rc;
  }
  CU_TRY(c, cudaStreamSynchronize(c->stream));
  CU_TRY(c, cudaGetLastError());
  return DNBC_OK;
}

extern "C" int dnbc_em_iterate_group(dnbc_ctx **ctxs, int n, int n_iters, double *loglik_out) {
  if (!ctxs || n < 1) return DNBC_EINVAL;
  int64_t total = 0;
  for (int i = 0; i < n; ++i) {
    if (!ctxs[i]) return DNBC_EINVAL;
    total += ctxs[i]->S;
  }
  if (total <= 0) return dnbc_fail(ctxs[0], DNBC_ESTATE, "em_iterate_group: no observations resident");
  for (int it = 0; it < n_iters; ++it) {
    for (int i = 0; i < n; ++i) {       // asynchronous launches: the GPUs run their shards concurrently
      int rc = dnbc_em_estep(ctxs[i], 0);
      if (rc) return rc;
    }
    int rc = dnbc_em_allreduce_group(ctxs, n);
    if (rc) return rc;
    if (loglik_out) {
      CU_TRY(ctxs[0], cudaSetDevice(ctxs[0]->device));
      CU_TRY(ctxs[0], cudaMemcpyAsync(loglik_out + it, ctxs[0]->d_stats, sizeof(double), cudaMemcpyDeviceToHost, ctxs[0]->stream));
    }
    for (int i = 0; i < n; ++i) {
      rc = dnbc_em_mstep(ctxs[i], total);
      if (rc) return rc;
    }
  }
  for (int i = 0; i < n; ++i) {
    CU_TRY(ctxs[i], cudaSetDevice(ctxs[i]->device));
    CU_TRY(ctxs[i], cudaStreamSynchronize(ctxs[i]->stream));
    CU_TRY(ctxs[i], cudaGetLastError());
  }
  return DNBC_OK;
}

// gather.h -- what job.cu needs from the NVLink gather (gather.cu): the per-round protocol of one rank.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/zoicb.h"
#include "kernels.h"

namespace zoicb {

int gather_device(const zoicb_gather* g);
int gather_rank(const zoicb_gather* g);
uint64_t gather_tile_rays(const zoicb_gather* g);   // records every rank contributes per round (at most)
// number of rounds all ranks run for these per-rank totals (counts[world], the same array on every rank)
uint64_t gather_rounds(const zoicb_gather* g, const uint64_t* counts);
// starts a job: remembers the per-rank totals and numbers its rounds; `st` is the stream the caller generates on;
// serial: shipping, waiting and consuming happen on `st` itself instead of the gather's own stream (no overlap)
cudaError_t gather_begin(zoicb_gather* g, const uint64_t* counts, cudaStream_t st, bool serial);
// where this rank's generate kernels write their records of `round`; `st` waits until that memory may be overwritten
cudaError_t gather_acquire(zoicb_gather* g, uint64_t round, cudaStream_t st, RayRecord** dst);
// this rank's m records of `round` are complete in stream order on `st`: ship / signal them; on the consumer rank also
// wait for every other rank's records of the round, run the consumer over them (totals into d_totals) and release the slot
cudaError_t gather_commit(zoicb_gather* g, uint64_t round, uint64_t m, cudaStream_t st, void* d_totals, int* launches);
// joins the gather's own streams into `st`
cudaError_t gather_end(zoicb_gather* g, cudaStream_t st);
bool gather_failed(zoicb_gather* g);   // a flag wait timed out since gather_begin

// job.cu: the consumer kernel (checksum + counts of a span of records)
cudaError_t launch_consume(const RayRecord* rays, uint64_t n, void* d_totals, cudaStream_t st, int* launches);

}  // namespace zoicb

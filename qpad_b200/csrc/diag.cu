// diag.cu -- device-resident staging of diagnostics, SURVEY.md §8(f) rank 3.
//
// The reference's writers take HOST arrays: field%write_hdf5 hands `f2(dim, 1:nr, 1:nzp)` of every plane to pwfield_pipe
// (fields/ufield_class.f03:214-244, hdf5io_class.f03:591-697), part3d%wr / part2d%wr hand every `dspl`-th particle, one attribute
// per dataset, to pwpart_3d_pipe / pwpart_2d_r (beam/part3d_class.f03:803-822, hdf5io_class.f03:1027-1190, :1220-1480: tnpp =
// int(npp / dspl), particles 1, 1 + dspl, ..., x3 shifted by z0 for a beam).  With the state living in HBM a dump must not stall
// the xi pipeline, so a `qpg_stage` owns a device buffer, a PINNED host buffer, a copy stream and an event:
//
//   qpg_stage_field / _part2d / _part3d   enqueue, on the CONTEXT's stream, one re-layout kernel that writes the datasets exactly
//                                         as the HDF5 writer wants them (contiguous per dataset), then make the stage's COPY stream
//                                         wait for it and start the device-to-host copy there -- the context's stream is free to
//                                         run the next 3D step while the copy engine drains the buffer;
//   qpg_stage_wait                        blocks the HOST until the copy has landed and returns the pinned pointer.
//
// Layouts of the staged buffer (fp64):
//   field   : [plane 0..P-1][comp 0..dim-1][slice 1..nzp][node 1..nr]      (one C-ordered (nzp, nr) dataset per plane and component)
//   part2d  : [0] = tnpp, then 6 datasets x1 x2 p1 p2 p3 q of `stride` entries each, the first tnpp valid
//   part3d  : [0] = tnpp, then 7 datasets x1 x2 x3+z0 p1 p2 p3 q          (stride = int(npp_hi / dspl) is returned to the caller)
//
// STATUS: the re-layout kernels and the copy / event protocol pass on the GPU (tests/test_gpu_extras.py, run by default since round 2)
// and, kernels only, in the host emulation of tests/emu (streams and events are no-ops there).
#include "common.cuh"

struct qpg_stage_s {
    qpg_ctx ctx;
    long cap;              // doubles
    double *dev, *host;    // host is pinned
    cudaStream_t copy;
    cudaEvent_t ready, done;
    long count;            // doubles of the transfer in flight
    bool busy;
};

// one thread per (slice, node): reads the P * dim contiguous values of the node, writes P * dim coalesced dataset rows
__global__ void __launch_bounds__(256) k_stage_field(const double *__restrict__ f2, double *__restrict__ out, int nr, int nzp, int P, int dim)
{
    const size_t n1 = (size_t)(nr + 2) * P * dim, plane = (size_t)nzp * nr;
    const long total = (long)nzp * nr;
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const int k = (int)(t / nr), j = (int)(t % nr) + 1;                 // slice k + 1, node j
        const double *src = f2 + (size_t)k * n1 + (size_t)j * P * dim;
        for (int pc = 0; pc < P * dim; pc++) out[(size_t)pc * plane + (size_t)t] = src[pc];
    }
}

struct StagePlanes { const double *a[7]; };
// hdf5io_class.f03:1061, :1126: tnpp = int(npp / dspl), entries 0, dspl, 2 dspl, ... (0-based); shift added to dataset `zplane`
__global__ void __launch_bounds__(256) k_stage_part(StagePlanes pl, int nplanes, const int *__restrict__ d_npp, int dspl, long stride, int zplane, double shift,
                                                    double *__restrict__ out)
{
    const int npp = *d_npp;
    long tnpp = npp / dspl;
    if (tnpp > stride) tnpp = stride;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = (double)tnpp;
    for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < tnpp; t += (long)gridDim.x * blockDim.x)
        for (int a = 0; a < nplanes; a++) {
            const double v = pl.a[a][t * dspl];
            out[1 + (size_t)a * stride + t] = a == zplane ? __dadd_rn(v, shift) : v;
        }
}

extern "C" int qpg_stage_create(qpg_stage *out, qpg_ctx ctx, long capacity_doubles)
{
    ARG_TRY(out && ctx && capacity_doubles > 0, "bad arg");
    qpg_stage s = new qpg_stage_s();
    memset(s, 0, sizeof(*s));
    s->ctx = ctx; s->cap = capacity_doubles;
    CUDA_TRY(cudaMalloc(&s->dev, sizeof(double) * capacity_doubles));
    CUDA_TRY(cudaMallocHost(&s->host, sizeof(double) * capacity_doubles));
    CUDA_TRY(cudaStreamCreateWithFlags(&s->copy, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&s->ready, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&s->done, cudaEventDisableTiming));
    *out = s;
    return 0;
}
extern "C" int qpg_stage_destroy(qpg_stage s)
{
    if (!s) return 0;
    cudaStreamSynchronize(s->copy);
    cudaEventDestroy(s->ready); cudaEventDestroy(s->done); cudaStreamDestroy(s->copy);
    cudaFree(s->dev); cudaFreeHost(s->host);
    delete s;
    return 0;
}
// the copy stream picks the buffer up behind the re-layout kernel; the context's stream does not wait for the copy
static int stage_ship(qpg_stage s, long count)
{
    qpg_ctx c = s->ctx;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(s->ready, c->stream));
    CUDA_TRY(cudaStreamWaitEvent(s->copy, s->ready, 0));
    CUDA_TRY(cudaMemcpyAsync(s->host, s->dev, sizeof(double) * count, cudaMemcpyDeviceToHost, s->copy));
    CUDA_TRY(cudaEventRecord(s->done, s->copy));
    s->count = count; s->busy = true;
    return 0;
}
#define STAGE_FREE(s) do { if ((s)->busy) { qpg_set_error("%s: a transfer is in flight on this stage (call qpg_stage_wait first)", __func__); return QPG_ERR_STATE; } } while (0)

extern "C" int qpg_stage_field(qpg_stage s, qpg_field f, long *count)
{
    ARG_TRY(s && f && f->has2d && f->ctx == s->ctx, "null handle / field without a volume / other context");
    STAGE_FREE(s);
    qpg_ctx c = s->ctx;
    const long n = (long)c->P * f->dim * f->nzp * c->nr;
    ARG_TRY(n <= s->cap, "stage buffer too small");
    long grid = ((long)f->nzp * c->nr + 255) / 256;
    if (grid > 148 * 8) grid = 148 * 8;
    k_stage_field<<<(int)grid, 256, 0, c->stream>>>(f->f2, s->dev, c->nr, f->nzp, c->P, f->dim);
    count_launch(c);
    if (count) *count = n;
    return stage_ship(s, n);
}
static int stage_part(qpg_stage s, const StagePlanes &pl, int nplanes, const int *d_npp, long npp_hi, int dspl, int zplane, double shift, long *stride_out)
{
    STAGE_FREE(s);
    qpg_ctx c = s->ctx;
    ARG_TRY(dspl >= 1, "dspl < 1");
    const long stride = npp_hi / dspl;
    const long n = 1 + (long)nplanes * stride;
    ARG_TRY(n <= s->cap, "stage buffer too small");
    long grid = (stride + 255) / 256;
    if (grid < 1) grid = 1;
    if (grid > 148 * 8) grid = 148 * 8;
    k_stage_part<<<(int)grid, 256, 0, c->stream>>>(pl, nplanes, d_npp, dspl, stride, zplane, shift, s->dev);
    count_launch(c);
    if (stride_out) *stride_out = stride;
    return stage_ship(s, n);
}
extern "C" int qpg_stage_part2d(qpg_stage s, qpg_part2d p, int dspl, long *stride)
{
    ARG_TRY(s && p && p->ctx == s->ctx, "null handle / other context");
    StagePlanes pl{{p->x1, p->x2, p->p1, p->p2, p->p3, p->q, nullptr}};
    return stage_part(s, pl, 6, p->d_npp, p->npp_hi < p->npmax ? p->npp_hi : p->npmax, dspl, -1, 0.0, stride);
}
extern "C" int qpg_stage_part3d(qpg_stage s, qpg_part3d p, int dspl, double z0, long *stride)
{
    ARG_TRY(s && p && p->ctx == s->ctx, "null handle / other context");
    StagePlanes pl{{p->x1, p->x2, p->x3, p->p1, p->p2, p->p3, p->q}};
    return stage_part(s, pl, 7, p->d_npp, p->npp_hi < p->npmax ? p->npp_hi : p->npmax, dspl, 2, z0, stride);
}
extern "C" int qpg_stage_wait(qpg_stage s, const double **host, long *count)
{
    ARG_TRY(s && host, "null arg");
    if (!s->busy) { qpg_set_error("qpg_stage_wait: nothing was staged"); return QPG_ERR_STATE; }
    CUDA_TRY(cudaEventSynchronize(s->done));
    s->busy = false;
    *host = s->host;
    if (count) *count = s->count;
    return 0;
}

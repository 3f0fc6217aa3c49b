// TEST INFRASTRUCTURE ONLY (see cuda_runtime.h in this directory): the translation unit of the host-emulation library.
// It provides what fields.cu / particles.cu provide in the real build (error reporting, the particle view, creation of contexts,
// fields and particle sets -- here plain host memory in the DEVICE layouts of DESIGN.md §3) and then includes the launch-rewritten
// device sources, so the extern "C" entry points of those files are the very code that runs on the GPU.
#include <cuda_runtime.h>
#include <cstdarg>
#include "common.cuh"

static thread_local char g_err[512] = "";
void qpg_set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap); }
int qpg_cuda_fail(cudaError_t e, const char *what) { qpg_set_error("CUDA error %d at %s", (int)e, what); return QPG_ERR_CUDA; }
extern "C" const char *qpg_last_error(void) { return g_err; }

// particles.cu:20-24
struct PartView {
    double *x1, *x2, *p1, *p2, *p3, *gamma, *psi, *q;
    const int *d_npp;
};
static PartView view_of(qpg_part2d p) { PartView v{p->x1, p->x2, p->p1, p->p2, p->p3, p->gamma, p->psi, p->q, p->d_npp}; return v; }

#ifdef EMU_HAVE_NEUTRAL
#include "neutral.cu.cpp"
#endif
#ifdef EMU_HAVE_SUBCYC
#include "subcyc.cu.cpp"
#endif
#ifdef EMU_HAVE_VPOT
#include "vpot.cu.cpp"
#endif
#ifdef EMU_HAVE_DIAG
#include "diag.cu.cpp"
#endif

extern "C" {
long emu_launches(void) { return emu::g_launches; }
long emu_barriers(void) { return emu::g_barriers; }

int emu_ctx_create(qpg_ctx *out, int nr, int max_mode, double dr, double dxi, int bnd)
{
    qpg_ctx c = new qpg_ctx_s();
    c->device = 0; c->stream = nullptr; c->own_stream = false;
    c->nr = nr; c->M = max_mode; c->P = 2 * max_mode + 1; c->dr = dr; c->dxi = dxi; c->bnd = bnd; c->relax = 0.0;
    c->launches = 0; c->tprof_on = false; c->capturing = false;
    *out = c;
    return 0;
}
int emu_ctx_destroy(qpg_ctx c) { delete c; return 0; }
long emu_ctx_launches(qpg_ctx c) { return c->launches; }

// field storage in the device layout: f1[(j*P + pl)*dim + c], f2 = nzp+1 such images
int emu_field_create(qpg_field *out, qpg_ctx c, int dim, int nzp, int has2d)
{
    qpg_field f = new qpg_field_s();
    f->ctx = c; f->dim = dim; f->nzp = nzp; f->has2d = has2d;
    f->n1 = (size_t)(c->nr + 2) * c->P * dim;
    f->f1 = (double *)calloc(f->n1, sizeof(double));
    f->f2 = has2d ? (double *)calloc(f->n1 * (nzp + 1), sizeof(double)) : nullptr;
    *out = f;
    return 0;
}
int emu_field_destroy(qpg_field f) { free(f->f1); free(f->f2); delete f; return 0; }
// host layout = the oracle's / the reference's: [plane][node 0..nr+1][comp] per slice (qpg_field_upload_f1 / _f2)
static void xfer(qpg_field f, double *host, double *image, int nslices, bool up)
{
    const int P = f->ctx->P, nn = f->ctx->nr + 2, dim = f->dim;
    for (int pl = 0; pl < P; pl++) for (int k = 0; k < nslices; k++) for (int j = 0; j < nn; j++) for (int c = 0; c < dim; c++) {
        double &h = host[(((size_t)pl * nslices + k) * nn + j) * dim + c], &d = image[(size_t)k * f->n1 + ((size_t)j * P + pl) * dim + c];
        if (up) d = h; else h = d;
    }
}
int emu_field_upload_f1(qpg_field f, const double *host) { xfer(f, const_cast<double *>(host), f->f1, 1, true); return 0; }
int emu_field_download_f1(qpg_field f, double *host) { xfer(f, host, f->f1, 1, false); return 0; }
int emu_field_upload_f2(qpg_field f, const double *host) { xfer(f, const_cast<double *>(host), f->f2, f->nzp + 1, true); return 0; }
int emu_field_download_f2(qpg_field f, double *host) { xfer(f, host, f->f2, f->nzp + 1, false); return 0; }

int emu_part2d_create(qpg_part2d *out, qpg_ctx c, double qbm, long npmax)
{
    qpg_part2d p = new qpg_part2d_s();
    memset(p, 0, sizeof(*p));
    p->ctx = c; p->qbm = qbm; p->npmax = npmax; p->npp_hi = 0;
    p->slab = (double *)calloc((size_t)8 * npmax, sizeof(double));
    double **pl[8] = {&p->x1, &p->x2, &p->p1, &p->p2, &p->p3, &p->gamma, &p->psi, &p->q};
    for (int a = 0; a < 8; a++) *pl[a] = p->slab + (size_t)a * npmax;
    p->d_npp = (int *)calloc(4, sizeof(int));
    *out = p;
    return 0;
}
int emu_part2d_destroy(qpg_part2d p) { free(p->slab); free(p->d_npp); delete p; return 0; }
long emu_part2d_npp(qpg_part2d p) { return p->d_npp[0]; }
// host AoS x(2,n) p(3,n) gamma psi q  <->  SoA planes (qpg_part2d_upload / _download)
int emu_part2d_upload(qpg_part2d p, const double *x, const double *pp, const double *g, const double *psi, const double *q, long n)
{
    for (long i = 0; i < n; i++) {
        p->x1[i] = x[2 * i]; p->x2[i] = x[2 * i + 1]; p->p1[i] = pp[3 * i]; p->p2[i] = pp[3 * i + 1]; p->p3[i] = pp[3 * i + 2];
        p->gamma[i] = g[i]; p->psi[i] = psi[i]; p->q[i] = q[i];
    }
    p->d_npp[0] = (int)n; p->npp_hi = n;
    return 0;
}
int emu_part2d_download(qpg_part2d p, double *x, double *pp, double *g, double *psi, double *q)
{
    for (long i = 0; i < p->d_npp[0]; i++) {
        x[2 * i] = p->x1[i]; x[2 * i + 1] = p->x2[i]; pp[3 * i] = p->p1[i]; pp[3 * i + 1] = p->p2[i]; pp[3 * i + 2] = p->p3[i];
        g[i] = p->gamma[i]; psi[i] = p->psi[i]; q[i] = p->q[i];
    }
    return 0;
}

int emu_part3d_create(qpg_part3d *out, qpg_ctx c, double qbm, double dt, long npmax)
{
    qpg_part3d p = new qpg_part3d_s();
    memset(p, 0, sizeof(*p));
    p->ctx = c; p->qbm = qbm; p->dt = dt; p->npmax = npmax;
    p->slab = (double *)calloc((size_t)7 * npmax, sizeof(double));
    double **pl[7] = {&p->x1, &p->x2, &p->x3, &p->p1, &p->p2, &p->p3, &p->q};
    for (int a = 0; a < 7; a++) *pl[a] = p->slab + (size_t)a * npmax;
    p->d_npp = (int *)calloc(4, sizeof(int));
    *out = p;
    return 0;
}
int emu_part3d_destroy(qpg_part3d p) { free(p->slab); free(p->d_npp); delete p; return 0; }
int emu_part3d_upload(qpg_part3d p, const double *x, const double *pp, const double *q, long n)
{
    for (long i = 0; i < n; i++) {
        p->x1[i] = x[3 * i]; p->x2[i] = x[3 * i + 1]; p->x3[i] = x[3 * i + 2]; p->p1[i] = pp[3 * i]; p->p2[i] = pp[3 * i + 1]; p->p3[i] = pp[3 * i + 2];
        p->q[i] = q[i];
    }
    p->d_npp[0] = (int)n; p->npp_hi = n;
    return 0;
}
}

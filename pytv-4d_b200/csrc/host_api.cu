// Host-buffer entry points: callers that hold numpy-style HOST arrays (the reference's default calling
// convention, tv_GPU.py:129-139: numpy in, numpy out) and do not manage device memory themselves.
#include <new>

#include "host_common.cuh"

using namespace pytvb;

namespace {

size_t elem_size(const pytvb_problem* pb) { return pb->dtype == PYTVB_F32 ? 4 : 8; }
size_t voxels(const pytvb_problem* pb) { return (size_t)pb->Nz * pb->M * pb->Ni * pb->Nj; }

struct DeviceBuf {
    void* p = nullptr;
    ~DeviceBuf() { if (p) cudaFree(p); }
    // zeroed: a reduce workspace starts with its arrival counter at 0 (finish_partials)
    cudaError_t alloc(size_t n) {
        cudaError_t e = cudaMalloc(&p, n ? n : 1);
        return e == cudaSuccess ? cudaMemset(p, 0, n ? n : 1) : e;
    }
};

// Device copies of the HOST arrays a problem descriptor points to - the (Ni,Nj) byte mask and the (Nz,M,Ni,Nj) time scale;
// dev problem = host problem with the pointers swapped.
int upload_mask(const pytvb_problem* host_pb, pytvb_problem* dev_pb, DeviceBuf* buf, DeviceBuf* ts_buf, cudaStream_t st) {
    *dev_pb = *host_pb;
    dev_pb->time_scale_lo = dev_pb->time_scale_hi = nullptr;      // whole volumes only
    if (host_pb->mask_static) {
        const size_t n = (size_t)host_pb->Ni * host_pb->Nj;
        PYTVB_CUDA(buf->alloc(n));
        PYTVB_CUDA(cudaMemcpyAsync(buf->p, host_pb->mask_static, n, cudaMemcpyHostToDevice, st));
        dev_pb->mask_static = (const uint8_t*)buf->p;
    }
    if (host_pb->time_scale) {
        const size_t n = voxels(host_pb) * elem_size(host_pb);
        PYTVB_CUDA(ts_buf->alloc(n));
        PYTVB_CUDA(cudaMemcpyAsync(ts_buf->p, host_pb->time_scale, n, cudaMemcpyHostToDevice, st));
        dev_pb->time_scale = ts_buf->p;
    }
    return PYTVB_OK;
}

}  // namespace

struct pytvb_cp_solver {
    pytvb_problem pb;          // device-side problem (mask_static on the device)
    double lam, sigma, tau, theta;
    DeviceBuf mask, tscale, x, xbar, x0, y, ws, scal;
    cudaStream_t st = nullptr;
    size_t img_bytes = 0, y_bytes = 0;
    ~pytvb_cp_solver() { if (st) cudaStreamDestroy(st); }
};

extern "C" {

int pytvb_tv_host(const pytvb_problem* pb, const void* x_host, void* G_host, void* norms_host_or_null, double* tv_out) {
    if (int rc = check_problem(pb)) return rc;
    PYTVB_REQUIRE(x_host && G_host && tv_out, "x_host, G_host and tv_out must not be NULL");
    PYTVB_REQUIRE(pb->z_offset == 0 && pb->Nz_global == pb->Nz, "pytvb_tv_host works on whole volumes");
    cudaStream_t st = nullptr;
    const size_t nb = voxels(pb) * elem_size(pb);
    DeviceBuf mask, tscale, x, G, norms, wsr, wst, dtv;
    pytvb_problem dpb;
    if (int rc = upload_mask(pb, &dpb, &mask, &tscale, st)) return rc;
    PYTVB_CUDA(x.alloc(nb));
    PYTVB_CUDA(G.alloc(nb));
    if (norms_host_or_null) PYTVB_CUDA(norms.alloc(nb));
    PYTVB_CUDA(wsr.alloc(pytvb_reduce_workspace_bytes(&dpb)));
    PYTVB_CUDA(wst.alloc(pytvb_tv_workspace_bytes(&dpb)));
    PYTVB_CUDA(dtv.alloc(sizeof(double)));
    PYTVB_CUDA(cudaMemcpyAsync(x.p, x_host, nb, cudaMemcpyHostToDevice, st));
    if (int rc = pytvb_tv(&dpb, x.p, G.p, norms_host_or_null ? norms.p : nullptr, (double*)dtv.p, nullptr, nullptr, wsr.p, wst.p, st)) return rc;
    PYTVB_CUDA(cudaMemcpyAsync(G_host, G.p, nb, cudaMemcpyDeviceToHost, st));
    if (norms_host_or_null) PYTVB_CUDA(cudaMemcpyAsync(norms_host_or_null, norms.p, nb, cudaMemcpyDeviceToHost, st));
    PYTVB_CUDA(cudaMemcpyAsync(tv_out, dtv.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    PYTVB_CUDA(cudaStreamSynchronize(st));
    return PYTVB_OK;
}

int pytvb_cp_create(const pytvb_problem* pb, double lam, double sigma, double tau, double theta, pytvb_cp_solver** out) {
    if (int rc = check_problem(pb)) return rc;
    PYTVB_REQUIRE(out, "out must not be NULL");
    PYTVB_REQUIRE(pb->z_offset == 0 && pb->Nz_global == pb->Nz, "the host-buffer solver works on whole volumes");
    PYTVB_REQUIRE(lam >= 0 && sigma > 0 && tau > 0, "lam >= 0, sigma > 0, tau > 0 required");
    pytvb_cp_solver* s = new (std::nothrow) pytvb_cp_solver();
    PYTVB_REQUIRE(s, "out of host memory");
    s->lam = lam; s->sigma = sigma; s->tau = tau; s->theta = theta;
    cudaError_t e = cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete s; set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e)); return PYTVB_ERR_CUDA; }
    int rc = upload_mask(pb, &s->pb, &s->mask, &s->tscale, s->st);
    s->img_bytes = voxels(pb) * elem_size(pb);
    s->y_bytes = s->img_bytes * (size_t)axes_of(pb).Nd;
    if (rc == PYTVB_OK) {
        e = s->x.alloc(s->img_bytes);
        if (e == cudaSuccess) e = s->xbar.alloc(s->img_bytes);
        if (e == cudaSuccess) e = s->x0.alloc(s->img_bytes);
        if (e == cudaSuccess) e = s->y.alloc(s->y_bytes);
        if (e == cudaSuccess) e = s->ws.alloc(pytvb_reduce_workspace_bytes(&s->pb));
        if (e == cudaSuccess) e = s->scal.alloc(2 * sizeof(double));
        if (e != cudaSuccess) { set_error("device allocation failed: %s", cudaGetErrorString(e)); rc = PYTVB_ERR_CUDA; }
    }
    if (rc != PYTVB_OK) { delete s; return rc; }
    *out = s;
    return PYTVB_OK;
}

int pytvb_cp_reset_host(pytvb_cp_solver* s, const void* x0_host) {
    PYTVB_REQUIRE(s && x0_host, "solver and x0_host must not be NULL");
    PYTVB_CUDA(cudaMemcpyAsync(s->x0.p, x0_host, s->img_bytes, cudaMemcpyHostToDevice, s->st));
    PYTVB_CUDA(cudaMemcpyAsync(s->x.p, s->x0.p, s->img_bytes, cudaMemcpyDeviceToDevice, s->st));
    PYTVB_CUDA(cudaMemcpyAsync(s->xbar.p, s->x0.p, s->img_bytes, cudaMemcpyDeviceToDevice, s->st));
    PYTVB_CUDA(cudaMemsetAsync(s->y.p, 0, s->y_bytes, s->st));
    PYTVB_CUDA(cudaStreamSynchronize(s->st));
    return PYTVB_OK;
}

int pytvb_cp_step_host(pytvb_cp_solver* s, const void* x0_host, void* x_host_or_null, double* energy_out) {
    PYTVB_REQUIRE(s && x0_host, "solver and x0_host must not be NULL");
    double* d_l21 = (double*)s->scal.p;
    double* d_fid = d_l21 + 1;
    PYTVB_CUDA(cudaMemcpyAsync(s->x0.p, x0_host, s->img_bytes, cudaMemcpyHostToDevice, s->st));
    if (int rc = pytvb_cp_dual(&s->pb, s->xbar.p, s->y.p, s->lam, s->sigma, d_l21, nullptr, nullptr, s->ws.p, s->st)) return rc;
    if (int rc = pytvb_cp_primal_rof(&s->pb, s->y.p, s->x.p, s->xbar.p, s->x0.p, s->tau, s->theta, d_fid, nullptr, nullptr, s->ws.p, s->st)) return rc;
    double h[2] = {0, 0};
    if (x_host_or_null) PYTVB_CUDA(cudaMemcpyAsync(x_host_or_null, s->x.p, s->img_bytes, cudaMemcpyDeviceToHost, s->st));
    PYTVB_CUDA(cudaMemcpyAsync(h, s->scal.p, sizeof(h), cudaMemcpyDeviceToHost, s->st));
    PYTVB_CUDA(cudaStreamSynchronize(s->st));
    if (energy_out) *energy_out = 0.5 * h[1] + s->lam * h[0];
    return PYTVB_OK;
}

int pytvb_cp_destroy(pytvb_cp_solver* s) {
    delete s;
    return PYTVB_OK;
}

}  // extern "C"

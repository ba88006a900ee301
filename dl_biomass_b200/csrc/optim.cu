// Adam over one flat parameter arena -- the optimiser step of /root/reference/main.py:84,172
// (torch.optim.Adam(model.parameters(), lr, weight_decay): L2 penalty folded into the gradient, bias-corrected
// moments, no amsgrad), as ONE launch over the concatenation of every parameter instead of a multi-tensor
// launch list: parameters, gradients and both moments live in four parallel fp32 buffers (dl_biomass_b200/optim.py).
// HBM-bound: 7 floats of traffic per parameter (read p, g, m, v; write p, m, v).
//
// The step counter is a DEVICE scalar (a CUDA-graph replay must advance it): every thread reads it, the last block to
// finish bumps it (ticket in state[1]).
#include <math.h>

#include "common.cuh"

namespace b2pn {

struct AdamParams {
    float *p;
    const float *g;
    float *m;
    float *v;
    int64_t n;        // multiple of 4
    float lr, beta1, beta2, eps, weight_decay, grad_scale;
    int64_t *state;   // [0] steps taken so far, [1] ticket
};

constexpr int ADAM_THREADS = 256;

__global__ void __launch_bounds__(ADAM_THREADS) adam_flat_kernel(const AdamParams a)
{
    const int64_t step = a.state[0] + 1;
    // bias corrections in double: 1 - beta^step loses everything in fp32 for beta2 = 0.999 at small steps
    const double bc1 = 1.0 - pow((double)a.beta1, (double)step);
    const double bc2 = 1.0 - pow((double)a.beta2, (double)step);
    const float step_size = (float)((double)a.lr / bc1);
    const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    const float b1 = a.beta1, b2 = a.beta2, wd = a.weight_decay, gs = a.grad_scale, eps = a.eps;
    const int64_t n4 = a.n >> 2;
    float4 *p4 = reinterpret_cast<float4 *>(a.p);
    const float4 *g4 = reinterpret_cast<const float4 *>(a.g);
    float4 *m4 = reinterpret_cast<float4 *>(a.m);
    float4 *v4 = reinterpret_cast<float4 *>(a.v);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 p = p4[i], m = m4[i], v = v4[i];
        const float4 g = g4[i];
        float pe[4] = {p.x, p.y, p.z, p.w}, me[4] = {m.x, m.y, m.z, m.w}, ve[4] = {v.x, v.y, v.z, v.w};
        const float ge[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float gr = fmaf(wd, pe[e], ge[e] * gs);          // grad + weight_decay * param
            me[e] = fmaf(1.f - b1, gr - me[e], me[e]);             // lerp, as torch
            ve[e] = fmaf(1.f - b2, gr * gr - ve[e], ve[e]);
            const float denom = sqrtf(ve[e]) * inv_sqrt_bc2 + eps;
            pe[e] -= step_size * (me[e] / denom);
        }
        p4[i] = make_float4(pe[0], pe[1], pe[2], pe[3]);
        m4[i] = make_float4(me[0], me[1], me[2], me[3]);
        v4[i] = make_float4(ve[0], ve[1], ve[2], ve[3]);
    }
    // every block has read state[0] before it takes a ticket; the last ticket holder publishes the new step count
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long t = atomicAdd(reinterpret_cast<unsigned long long *>(a.state + 1), 1ull);
        if (t == (unsigned long long)gridDim.x - 1ull) {
            a.state[1] = 0;
            a.state[0] = step;
            __threadfence();
        }
    }
}

}  // namespace b2pn

extern "C" int b2pn_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr,
                              float beta1, float beta2, float eps, float weight_decay, float grad_scale, int64_t *state,
                              b2pn_stream_t stream)
{
    if (n < 0 || (n & 3)) return B2PN_EINVAL;
    if (n == 0) return B2PN_OK;
    if (!param || !grad || !exp_avg || !exp_avg_sq || !state) return B2PN_EINVAL;
    if ((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15u) != 0) return B2PN_EINVAL;
    b2pn::AdamParams a = {param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, grad_scale, state};
    int sms = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    int64_t blocks = (n / 4 + b2pn::ADAM_THREADS - 1) / b2pn::ADAM_THREADS;
    const int64_t cap = (int64_t)sms * 8;  // grid-stride beyond 8 blocks per SM
    if (blocks > cap) blocks = cap;
    b2pn::adam_flat_kernel<<<(unsigned)blocks, b2pn::ADAM_THREADS, 0, (cudaStream_t)stream>>>(a);
    b2pn::note_launch();
    B2PN_LAUNCH_CHECK();
    return B2PN_OK;
}

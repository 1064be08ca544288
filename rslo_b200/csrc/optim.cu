// Optimizer step of the training loop in two launches: global gradient norm, then clip + weight decay + Adam.
//
// Replaces, for the path's 12.0 M fp32 parameters (SURVEY §8 f-N2):
//   torch.nn.utils.clip_grad_norm_(net.parameters(), 10.0)                       train_hdf5.py:671
//   OptimWrapper.step(): p.mul_(1 - wd*lr) for every trainable parameter (true_wd, bn_wd), then
//   torch.optim.Adam(betas=(mom, 0.99), eps=1e-8).step()                         rslo/torchplus/train/fastai_optim.py:181-194,
//                                                                                rslo/builder/optimizer_builder.py:101-118
// which in the reference is ~300 parameters x (norm, mul, 2 x mul/add, addcmul, sqrt, div, addcdiv) small launches.
//
// Layout: the gradients already live in ONE flat fp32 buffer (rslo_b200/utils/distributed.py, the all-reduce
// bucket); the Adam moments are two more flat buffers with the same offsets.  Parameters stay where the
// state_dict has them: a chunk table {parameter pointer, flat offset, count, flags} maps each CTA to at most
// ADAM_CHUNK contiguous elements of one parameter.  Both kernels are pure streaming (HBM-bound: 4 B/element for
// the norm, 28 B/element for the step).
//
// The squared norm is reduced in double, partial sums per CTA are combined by the last CTA to arrive in index
// order -> the same bits on every run and on every rank (the clip coefficient multiplies every gradient).
#include "common.cuh"

namespace rslo {
namespace {

constexpr int NORM_THREADS = 256;
constexpr int NORM_MAX_CTAS = 1184;          // 148 SMs x 8

__global__ void __launch_bounds__(NORM_THREADS)
k_grad_sumsq(const float* __restrict__ g, size_t n, double* __restrict__ partial, unsigned int* __restrict__ counter,
             double* __restrict__ out)
{
    __shared__ double red[NORM_THREADS / 32];
    __shared__ bool last;
    double acc = 0.0;
    const size_t n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(g4 + i);
        // fp32 products are exact in double; four of them are summed in double as well
        acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {
        const float v = __ldg(g + (n4 << 2) + threadIdx.x);
        acc += (double)v * v;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < NORM_THREADS / 32; ++w) s += red[w];
        partial[blockIdx.x] = s;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double s = 0.0;
        for (unsigned int i = 0; i < gridDim.x; ++i) s += ((volatile double*)partial)[i];
        *out = s;
        *counter = 0;                          // left zeroed for the next call
    }
}

constexpr int ADAM_THREADS = 256;

struct AdamHyper {
    float grad_scale, max_norm, lr, beta1, beta2, eps, decay, l2, inv_bc1, inv_sqrt_bc2;
};

__global__ void __launch_bounds__(ADAM_THREADS)
k_adam_step(const rslo_adam_chunk_t* __restrict__ chunks, float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, const double* __restrict__ sumsq, AdamHyper h, int write_grad)
{
    const rslo_adam_chunk_t c = chunks[blockIdx.x];
    // clip_grad_norm_: total_norm = ||grad_scale * g||_2, coef = max_norm / (total_norm + 1e-6), applied when < 1
    float coef = h.grad_scale;
    if (sumsq != nullptr && h.max_norm > 0.f) {
        const float total = (float)(sqrt(*sumsq) * (double)h.grad_scale);
        const float cc = h.max_norm / (total + 1e-6f);
        if (cc < 1.f) coef *= cc;
    }
    float* __restrict__ p = c.p;
    const size_t base = c.off;
    const float step = h.lr * h.inv_bc1;
    for (unsigned int i = threadIdx.x; i < c.n; i += ADAM_THREADS) {
        float w = p[i] * h.decay;                               // true weight decay: p *= 1 - wd*lr  (1 when unused)
        if (c.flags & 1u) {                                     // parameter received a gradient this step
            float gi = g[base + i] * coef;
            if (write_grad) g[base + i] = gi;
            gi = fmaf(h.l2, w, gi);                             // classic L2 (Adam's weight_decay), 0 with true_wd
            const float mi = fmaf(1.f - h.beta1, gi - m[base + i], m[base + i]);             // exp_avg.lerp_(g, 1-b1)
            const float vi = fmaf(1.f - h.beta2, gi * gi, h.beta2 * v[base + i]);            // mul_(b2).addcmul_(g,g,1-b2)
            m[base + i] = mi;
            v[base + i] = vi;
            const float denom = sqrtf(vi) * h.inv_sqrt_bc2 + h.eps;
            w -= step * (mi / denom);
        }
        p[i] = w;
    }
}

}  // namespace
}  // namespace rslo

using namespace rslo;

extern "C" size_t rslo_grad_norm_workspace_bytes(void) { return (size_t)NORM_MAX_CTAS * sizeof(double) + 256; }

extern "C" int rslo_grad_sumsq(const float* grad, size_t n, double* sumsq_out, void* workspace, size_t workspace_bytes,
                               rslo_stream_t stream)
{
    if (workspace == nullptr || workspace_bytes < rslo_grad_norm_workspace_bytes() || ((uintptr_t)grad & 15)) {
        set_last_error("rslo_grad_sumsq: workspace too small or gradient buffer not 16-byte aligned", cudaErrorInvalidValue);
        return (int)cudaErrorInvalidValue;
    }
    unsigned int* counter = (unsigned int*)workspace;               // zeroed once by the caller, left zeroed
    double* partial = (double*)((char*)workspace + 256);
    int ctas = cdiv((long long)(n >> 2), NORM_THREADS * 4);
    if (ctas < 1) ctas = 1;
    if (ctas > NORM_MAX_CTAS) ctas = NORM_MAX_CTAS;
    RSLO_COUNT();
    k_grad_sumsq<<<ctas, NORM_THREADS, 0, (cudaStream_t)stream>>>(grad, n, partial, counter, sumsq_out);
    RSLO_CHECK_LAUNCH("rslo_grad_sumsq");
    return 0;
}

extern "C" int rslo_adam_step(const rslo_adam_chunk_t* chunks_dev, int n_chunks, float* grad, float* exp_avg,
                              float* exp_avg_sq, const double* sumsq, float grad_scale, float max_norm, float lr,
                              float beta1, float beta2, float eps, float weight_decay, int true_wd, int step,
                              int write_clipped_grad, rslo_stream_t stream)
{
    if (n_chunks <= 0) return 0;
    if (step < 1) {
        set_last_error("rslo_adam_step: step counts from 1", cudaErrorInvalidValue);
        return (int)cudaErrorInvalidValue;
    }
    AdamHyper h;
    h.grad_scale = grad_scale;
    h.max_norm = max_norm;
    h.lr = lr;
    h.beta1 = beta1;
    h.beta2 = beta2;
    h.eps = eps;
    h.decay = true_wd ? 1.f - weight_decay * lr : 1.f;
    h.l2 = true_wd ? 0.f : weight_decay;
    // bias corrections in double on the host, as torch computes them in python floats
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    h.inv_bc1 = (float)(1.0 / bc1);
    h.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    RSLO_COUNT();
    k_adam_step<<<n_chunks, ADAM_THREADS, 0, (cudaStream_t)stream>>>(chunks_dev, grad, exp_avg, exp_avg_sq, sumsq, h,
                                                                    write_clipped_grad);
    RSLO_CHECK_LAUNCH("rslo_adam_step");
    return 0;
}

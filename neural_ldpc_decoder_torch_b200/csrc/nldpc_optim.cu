// nldpc_optim.cu — the tail of one training step on the flat weight vector, in ONE launch
// (/root/reference/train/train_BoostedNeuralLDPCDecoder.py:291-294):
//     torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)      :291
//     optimizer.step()            (torch.optim.Adam, defaults)          :292-293
//     model._apply_constraints()  (clamp into allowed_weight_range)     BoostedNeuralLDPCDecoder.py:153-179
// The decoders' parameter vectors are tiny (20 ... T*E = a few thousand floats), so the reference's ~15 small launches per
// step (norms, stack, clamp, mul, Adam foreach ops, T clamps) are pure launch latency; here one CTA does all of it and the
// step counter lives on the device, which makes the launch CUDA-graph replayable.
#include <cuda_runtime.h>
#include <stdint.h>

namespace nldpc {

constexpr int kOptThreads = 256;

// state[0] = step count (as float, exact up to 2^24 steps), state[1] = last total gradient norm (before clipping)
__global__ void __launch_bounds__(kOptThreads) clip_adam_clamp_kernel(float *__restrict__ p, float *__restrict__ g, float *__restrict__ m,
                                                                       float *__restrict__ v, float *__restrict__ state, int n,
                                                                       int n_norm, float grad_scale, float max_norm, double lr,
                                                                       const float *__restrict__ lr_dev, double beta1,
                                                                       double beta2, double eps_d, float lo, float hi) {
    __shared__ double red[kOptThreads / 32];
    __shared__ float s_coef;
    double acc = 0.0;
    // clip_grad_norm_(model.parameters()): the norm runs over ALL n_norm gradients, of which only the first n belong to
    // parameters the optimiser updates (get_trainable_parameters(); the rest only receive the clipped gradient)
    for (int i = threadIdx.x; i < n_norm; i += kOptThreads) {
        const float gi = g[i] * grad_scale;
        acc += (double)gi * (double)gi;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < kOptThreads / 32; w++) tot += red[w];
        const float norm = (float)sqrt(tot);
        float coef = 1.0f;
        if (max_norm > 0.0f) {                                   // clip_grad_norm_: coef = min(max_norm / (norm + 1e-6), 1)
            coef = max_norm / (norm + 1e-6f);
            coef = coef > 1.0f ? 1.0f : coef;
        }
        s_coef = coef;
        state[1] = norm;
    }
    __syncthreads();
    const float coef = s_coef * grad_scale;
    const float step = state[0] + 1.0f;
    // torch.optim.Adam (single-tensor formulation): bias corrections from the step count
    // (computed in double like the Python scalars of torch's implementation, then used as fp32 scalars)
    const float bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
    if (lr_dev) lr = (double)*lr_dev;          // learning rate read at run time: a replayed CUDA graph follows the schedule
    const float step_size = (float)(lr / (1.0 - pow(beta1, (double)step)));
    const float w1 = (float)(1.0 - beta1), b2 = (float)beta2, w2 = (float)(1.0 - beta2), eps = (float)eps_d;
    for (int i = threadIdx.x; i < n; i += kOptThreads) {
        const float gi = g[i] * coef;
        const float mi = m[i] + (gi - m[i]) * w1;            // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = v[i] * b2 + w2 * gi * gi;        // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        float pi = p[i] - step_size * (mi / denom);                      // param.addcdiv_(exp_avg, denom, -step_size)
        pi = fminf(fmaxf(pi, lo), hi);                                   // _apply_constraints
        g[i] = gi;                                                       // the clipped gradient, as clip_grad_norm_ leaves it
        m[i] = mi;
        v[i] = vi;
        p[i] = pi;
    }
    for (int i = n + threadIdx.x; i < n_norm; i += kOptThreads) g[i] = g[i] * coef;
    __syncthreads();
    if (threadIdx.x == 0) state[0] = step;
}

int launch_clip_adam_clamp(float *p, float *g, float *m, float *v, float *state, int n, int n_norm, float grad_scale, float max_norm,
                           double lr, const float *lr_dev, double beta1, double beta2, double eps, float lo, float hi, cudaStream_t st) {
    clip_adam_clamp_kernel<<<1, kOptThreads, 0, st>>>(p, g, m, v, state, n, n_norm, grad_scale, max_norm, lr, lr_dev, beta1, beta2, eps,
                                                      lo, hi);
    return (int)cudaGetLastError();
}

}  // namespace nldpc

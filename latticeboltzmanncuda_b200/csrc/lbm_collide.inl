// lbm_collide.inl -- the two collision operators, included inside the anonymous namespace of every
// translation unit that steps the lattice (lbm_kernels.cu, lbm_inplace.cu).

// ---- collision ---------------------------------------------------------------------------

// Reference arithmetic (SimCalcFuncs.cu:100-155): fp32 moments accumulated in index order,
// equilibrium promoted to double by the double literals, narrowed to float, float subtract,
// float divide by the timescale.  Expression shapes are kept so that nvcc contracts them the
// way it contracts the reference's.
template <int Q, int MOMENT>
__device__ __forceinline__ void collide_fp64(float (&f)[Q], float timescale)
{
    float density = 0.0f, ux = 0.0f, uy = 0.0f, uz = 0.0f;
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const float s = (MOMENT == 0) ? dir_split(i) : 1.0f;
        density += f[i];
        if (dir_c(i, 0) != 0) ux += f[i] * (float)dir_c(i, 0) * s;
        if (dir_c(i, 1) != 0) uy += f[i] * (float)dir_c(i, 1) * s;
        if (dir_c(i, 2) != 0) uz += f[i] * (float)dir_c(i, 2) * s;
    }
    if (density != 0.0) {
        ux /= density; uy /= density; uz /= density;
    }
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        float common_term = (((float)dir_c(i, 0) * ux) + ((float)dir_c(i, 1) * uy) + ((float)dir_c(i, 2) * uz));
        float eq = dir_weight<Q>(i) * density *
                   (1 + (3.0 * common_term) + (4.5 * common_term * common_term) - (1.5 * ((ux * ux) + (uy * uy) + (uz * uz))));
        f[i] = f[i] + (eq - f[i]) / timescale;
    }
}

// Throughput arithmetic: all fp32, opposite directions share the even part of the
// equilibrium, relaxation by multiplication with 1/tau.
template <int Q, int MOMENT>
__device__ __forceinline__ void collide_fp32(float (&f)[Q], float inv_tau)
{
    float rho = f[0], mx = 0.0f, my = 0.0f, mz = 0.0f;
#pragma unroll
    for (int i = 1; i < Q; ++i) {
        const float s = (MOMENT == 0) ? dir_split(i) : 1.0f;
        rho += f[i];
        if (dir_c(i, 0) != 0) mx = fmaf(f[i], (float)dir_c(i, 0) * s, mx);
        if (dir_c(i, 1) != 0) my = fmaf(f[i], (float)dir_c(i, 1) * s, my);
        if (dir_c(i, 2) != 0) mz = fmaf(f[i], (float)dir_c(i, 2) * s, mz);
    }
    const float inv_rho = (rho != 0.0f) ? 1.0f / rho : 1.0f;
    const float ux = mx * inv_rho, uy = my * inv_rho, uz = mz * inv_rho;
    const float base = fmaf(-1.5f, fmaf(ux, ux, fmaf(uy, uy, uz * uz)), 1.0f);
    float moving = 0.0f;                               // sum of the equilibria of all moving populations
#pragma unroll
    for (int i = 1; i < Q; ++i) {
        if (dir_inv(i) < i) continue;                  // handle each +/- pair once
        const int j = dir_inv(i);
        const float cu = (float)dir_c(i, 0) * ux + (float)dir_c(i, 1) * uy + (float)dir_c(i, 2) * uz;
        const float wr = (float)dir_weight<Q>(i) * rho;
        const float even = wr * fmaf(4.5f * cu, cu, base);
        const float odd = wr * 3.0f * cu;
        moving = fmaf(2.0f, even, moving);
        f[i] = fmaf(inv_tau, (even + odd) - f[i], f[i]);
        f[j] = fmaf(inv_tau, (even - odd) - f[j], f[j]);
    }
    // the rest equilibrium is what is left of rho: sum_i feq_i == rho holds to round-off instead of
    // drifting with the rounding of the fp32 weights (their sum is not exactly 1)
    f[0] = fmaf(inv_tau, (rho - moving) - f[0], f[0]);
}


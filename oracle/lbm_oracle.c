/*
 * lbm_oracle.c -- CPU restatement of the reference's per-timestep kernels.
 * TEST INFRASTRUCTURE ONLY (see lbm_oracle.h).  Citations relative to /root/reference.
 *
 * Written as plain loops over the reference's own AoS record so that each statement
 * can be held against the kernel line it restates.  OpenMP over x-slices is used only
 * to make the timed CPU baseline use the host's cores; results do not depend on the
 * thread count except for lbo_reflect's force sum, which is kept serial.
 */
#include "lbm_oracle.h"

#include <string.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* SimCalcFuncs.cu:12 */
const int LBO_DIRECTIONS[27][3] = {
    { 0, 0, 0},
    {-1, 0, 0}, { 0,-1, 0}, { 0, 0,-1}, { 0, 0, 1}, { 0, 1, 0}, { 1, 0, 0},
    {-1,-1, 0}, {-1, 0,-1}, {-1, 0, 1}, {-1, 1, 0}, { 0,-1,-1}, { 0,-1, 1},
    { 0, 1,-1}, { 0, 1, 1}, { 1,-1, 0}, { 1, 0,-1}, { 1, 0, 1}, { 1, 1, 0},
    {-1,-1,-1}, {-1,-1, 1}, {-1, 1,-1}, {-1, 1, 1}, { 1,-1,-1}, { 1,-1, 1}, { 1, 1,-1}, { 1, 1, 1}
};
/* SimCalcFuncs.cu:13 */
const int LBO_INVERSE[27] = {0, 6, 5, 4, 3, 2, 1, 18, 17, 16, 15, 14, 13, 12, 11, 10, 9, 8, 7,
                             26, 25, 24, 23, 22, 21, 20, 19};
/* SimCalcFuncs.cu:14 -- the literals are doubles narrowed to float at initialisation */
const float LBO_SPLIT[27] = {0.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0,
                             0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5,
                             1.0/3.0, 1.0/3.0, 1.0/3.0, 1.0/3.0, 1.0/3.0, 1.0/3.0, 1.0/3.0, 1.0/3.0};

void lbo_set_num_threads(int n)
{
    omp_set_num_threads(n > 0 ? n : 1);
}

int lbo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* get_lattice_at_coords' domain test, SimCalcFuncs.cu:37-40 */
static inline int in_domain(int x, int y, int z, int X, int Y, int Z)
{
    return !(x < 0 || y < 0 || z < 0 || x >= X || y >= Y || z >= Z);
}

void lbo_prime_points(lbo_lattice_point *cells, int X, int Y, int Z)
{
    for (int x = 0; x < X; ++x)
        for (int y = 0; y < Y; ++y)
            for (int z = 0; z < Z; ++z) {
                lbo_lattice_point *c = &cells[lbo_index(x, y, z, Y, Z)];
                c->x = x; c->y = y; c->z = z;
            }
}

void lbo_stream(const lbo_lattice_point *src, lbo_lattice_point *dst, int X, int Y, int Z, int bc)
{
    size_t n = (size_t)X * Y * Z;
    /* Lattice.cpp:126-132: the new lattice starts as a byte copy of the old one */
    if (dst != src) memcpy(dst, src, n * sizeof(*src));
#pragma omp parallel for schedule(static)
    for (int x = 0; x < X; ++x)
        for (int y = 0; y < Y; ++y)
            for (int z = 0; z < Z; ++z) {
                lbo_lattice_point *c = &dst[lbo_index(x, y, z, Y, Z)];
                /* get_lattice_point(templattice, true), SimCalcFuncs.cu:81 */
                c->x = x; c->y = y; c->z = z;
                for (int i = 0; i < 27; ++i) {
                    int nx = x + LBO_DIRECTIONS[i][0];
                    int ny = y + LBO_DIRECTIONS[i][1];
                    int nz = z + LBO_DIRECTIONS[i][2];
                    int j = LBO_INVERSE[i];
                    if (bc == LBO_BC_PERIODIC) {
                        nx = (nx + X) % X; ny = (ny + Y) % Y; nz = (nz + Z) % Z;
                    }
                    if (in_domain(nx, ny, nz, X, Y, Z))
                        c->f[j] = src[lbo_index(nx, ny, nz, Y, Z)].f[j];   /* :89 */
                    else
                        c->f[j] = 0.0;                                     /* :93 */
                }
            }
}

/* one cell of calculate_collision, SimCalcFuncs.cu:100-155.  Expression shapes follow the
 * kernel: float moments in index order, equilibrium promoted to double by the double
 * literals, narrowed to float, then a float subtract and a float divide by timescale. */
static inline void collide_cell_ref(float *f, float timescale)
{
    float density = 0.0;
    float ux = 0.0, uy = 0.0, uz = 0.0;
    for (int i = 0; i < 27; ++i) {
        density += f[i];
        ux += f[i] * LBO_DIRECTIONS[i][0] * LBO_SPLIT[i];
        uy += f[i] * LBO_DIRECTIONS[i][1] * LBO_SPLIT[i];
        uz += f[i] * LBO_DIRECTIONS[i][2] * LBO_SPLIT[i];
    }
    if (density != 0.0) {
        ux /= density; uy /= density; uz /= density;
    }
    for (int i = 0; i < 27; ++i) {
        double w = (i == 0) ? (8.0/27.0) : (i < 7) ? (2.0/27.0) : (i < 19) ? (1.0/54.0) : (1.0/216.0);
        float common_term = ((LBO_DIRECTIONS[i][0] * ux) + (LBO_DIRECTIONS[i][1] * uy) + (LBO_DIRECTIONS[i][2] * uz));
        float eq = w * density * (1 + (3.0 * common_term) + (4.5 * common_term * common_term)
                                  - (1.5 * ((ux * ux) + (uy * uy) + (uz * uz))));
        f[i] = f[i] + (eq - f[i]) / timescale;     /* atomicAdd on the thread's own cell */
    }
}

/* build-extension collide: q in {19,27}, standard or SPLIT moments, fp64 or fp32 equilibrium */
static inline void collide_cell_ext(float *f, int q, float tau, int moment, int eq64)
{
    float rho = 0.0f, ux = 0.0f, uy = 0.0f, uz = 0.0f;
    for (int i = 0; i < q; ++i) {
        float s = (moment == LBO_MOMENT_REFERENCE) ? LBO_SPLIT[i] : 1.0f;
        rho += f[i];
        ux += f[i] * LBO_DIRECTIONS[i][0] * s;
        uy += f[i] * LBO_DIRECTIONS[i][1] * s;
        uz += f[i] * LBO_DIRECTIONS[i][2] * s;
    }
    if (rho != 0.0f) { ux /= rho; uy /= rho; uz /= rho; }
    for (int i = 0; i < q; ++i) {
        double w;
        if (q == 27) w = (i == 0) ? (8.0/27.0) : (i < 7) ? (2.0/27.0) : (i < 19) ? (1.0/54.0) : (1.0/216.0);
        else         w = (i == 0) ? (1.0/3.0)  : (i < 7) ? (1.0/18.0) : (1.0/36.0);
        float cu = (LBO_DIRECTIONS[i][0] * ux) + (LBO_DIRECTIONS[i][1] * uy) + (LBO_DIRECTIONS[i][2] * uz);
        float eq;
        if (eq64) {
            eq = (float)(w * rho * (1 + (3.0 * cu) + (4.5 * cu * cu) - (1.5 * ((ux * ux) + (uy * uy) + (uz * uz)))));
            f[i] = f[i] + (eq - f[i]) / tau;
        } else {
            float usq = ux * ux + uy * uy + uz * uz;
            eq = (float)w * rho * ((1.0f - 1.5f * usq) + 3.0f * cu + 4.5f * cu * cu);
            f[i] = f[i] + (1.0f / tau) * (eq - f[i]);
        }
    }
}

void lbo_collide(lbo_lattice_point *cells, int X, int Y, int Z, float tau, int q, int moment)
{
    long n = (long)X * Y * Z;
#pragma omp parallel for schedule(static)
    for (long k = 0; k < n; ++k) {
        if (q == 27 && moment == LBO_MOMENT_REFERENCE) collide_cell_ref(cells[k].f, tau);
        else collide_cell_ext(cells[k].f, q, tau, moment, 1);
    }
}

int lbo_get_reflection_direction(const lbo_reflection_values *t, int slot)
{
    /* SimCalcFuncs.cu:18-30 / Lattice.cpp:465-477: slot s lives in int s/6 at bit 5*(s%6) */
    unsigned word = (unsigned)t->reflection_directions[slot / 6];
    return (int)((word >> (5 * (slot % 6))) & 0x1Fu);
}

void lbo_set_reflection_direction(lbo_reflection_values *t, int slot, unsigned value)
{
    /* Lattice.cpp:450-463 */
    unsigned shift = 5u * (unsigned)(slot % 6);
    unsigned word = (unsigned)t->reflection_directions[slot / 6];
    word &= ~(0x1Fu << shift);
    word |= value << shift;
    t->reflection_directions[slot / 6] = (int32_t)word;
}

void lbo_reflect(const lbo_lattice_point *src, lbo_lattice_point *dst, int X, int Y, int Z,
                 const lbo_reflection_values *tables, const int32_t *table_index, float *force)
{
    size_t n = (size_t)X * Y * Z;
    if (dst != src) memcpy(dst, src, n * sizeof(*src));   /* Lattice.cpp:176-182 */

    /* scatter part, SimCalcFuncs.cu:162-187 (serial: the reference's atomics make the
     * per-target sums order-dependent in fp32; the oracle fixes x,y,z,slot order) */
    for (int x = 0; x < X; ++x)
        for (int y = 0; y < Y; ++y)
            for (int z = 0; z < Z; ++z) {
                size_t k = lbo_index(x, y, z, Y, Z);
                const lbo_lattice_point *cur = &src[k];
                if (!cur->is_reflected || !tables || !table_index || table_index[k] < 0) continue;
                const lbo_reflection_values *t = &tables[table_index[k]];
                for (int s = 1; s < 81; ++s) {
                    int d = lbo_get_reflection_direction(t, s);
                    if (d == 0) continue;
                    int nx = x + LBO_DIRECTIONS[d][0];
                    int ny = y + LBO_DIRECTIONS[d][1];
                    int nz = z + LBO_DIRECTIONS[d][2];
                    if (!in_domain(nx, ny, nz, X, Y, Z)) continue;
                    float moved = cur->f[s / 3] * t->reflection_weight[s];
                    dst[lbo_index(nx, ny, nz, Y, Z)].f[d] += moved;           /* :177 */
                    if (force && s < 27) {
                        /* :181-183 -- incoming direction indexed by the SLOT s, not s/3 */
                        force[0] += moved * ((LBO_DIRECTIONS[d][0] * LBO_SPLIT[d]) - (LBO_DIRECTIONS[s][0] * LBO_SPLIT[s]));
                        force[1] += moved * ((LBO_DIRECTIONS[d][1] * LBO_SPLIT[d]) - (LBO_DIRECTIONS[s][1] * LBO_SPLIT[s]));
                        force[2] += moved * ((LBO_DIRECTIONS[d][2] * LBO_SPLIT[d]) - (LBO_DIRECTIONS[s][2] * LBO_SPLIT[s]));
                    }
                }
            }
    /* zeroing part, SimCalcFuncs.cu:189-197 */
#pragma omp parallel for schedule(static)
    for (long k = 0; k < (long)n; ++k)
        if (src[k].is_internal)
            for (int i = 0; i < 27; ++i) dst[k].f[i] = 0.0;
}

void lbo_source(lbo_lattice_point *cells, int X, int Y, int Z, int kind, const float *flow27)
{
    if (kind == LBO_SRC_NONE || !flow27) return;
#pragma omp parallel for schedule(static)
    for (int x = 0; x < X; ++x)
        for (int y = 0; y < Y; ++y)
            for (int z = 0; z < Z; ++z) {
                int hit = (kind == LBO_SRC_CELL_111) ? (x == 1 && y == 1 && z == 1)   /* FlowCriterion.cu:5-8 */
                        : (kind == LBO_SRC_LATTICE_PATTERN) ? ((x + 2 * y + 3 * z) % 5 == 0 && x >= 1)   /* ours: not a box */
                                                     : (x == 0);                       /* FlowCriterion.cu:10-13 */
                if (!hit) continue;
                float *f = cells[lbo_index(x, y, z, Y, Z)].f;
                for (int i = 0; i < 27; ++i) f[i] += flow27[i];                        /* SimCalcFuncs.cu:210-213 */
            }
}

void lbo_step_cli(lbo_lattice_point *cells, lbo_lattice_point *scratch, int X, int Y, int Z,
                  float tau, const lbo_reflection_values *tables, const int32_t *table_index,
                  float *force, int src_kind, const float *flow27)
{
    /* LatticeBoltzmannCuda.cpp:89-92 */
    lbo_stream(cells, scratch, X, Y, Z, LBO_BC_ZERO_INFLOW);
    lbo_reflect(scratch, cells, X, Y, Z, tables, table_index, force);
    lbo_collide(cells, X, Y, Z, tau, 27, LBO_MOMENT_REFERENCE);
    lbo_source(cells, X, Y, Z, src_kind, flow27);
}

void lbo_step_lattice(lbo_lattice_point *cells, lbo_lattice_point *scratch, int X, int Y, int Z, float tau)
{
    /* Lattice.cpp:222-226 */
    size_t n = (size_t)X * Y * Z;
    lbo_stream(cells, scratch, X, Y, Z, LBO_BC_ZERO_INFLOW);
    memcpy(cells, scratch, n * sizeof(*cells));
    lbo_collide(cells, X, Y, Z, tau, 27, LBO_MOMENT_REFERENCE);
}

unsigned lbo_compress_ints(const int *array, int num, int bit_length)
{
    /* Utils.cpp:4-29: first element ends up in the highest bits; the shift is always 5 */
    unsigned limit = 1u << bit_length;
    unsigned out = 0;
    for (int i = 0; i < num; ++i) {
        if ((unsigned)array[i] < limit) {
            out |= (unsigned)array[i];
            if (i < num - 1) out <<= 5;
        }
    }
    return out;
}

/* ---- build-extension SoA stepper (throughput configurations) --------------------- */

static inline double q_weight(int q, int i)
{
    if (q == 27) return (i == 0) ? (8.0/27.0) : (i < 7) ? (2.0/27.0) : (i < 19) ? (1.0/54.0) : (1.0/216.0);
    return (i == 0) ? (1.0/3.0) : (i < 7) ? (1.0/18.0) : (1.0/36.0);
}

void lbo_step_soa(const float *fin, float *fout, int X, int Y, int Z, int q, float tau,
                  const int bc[3], const float wall_u[6][3], const uint8_t *solid,
                  int moment, int eq64, double *force)
{
    const size_t n = (size_t)X * Y * Z;
    const int dims[3] = {X, Y, Z};
    double Fx = 0.0, Fy = 0.0, Fz = 0.0;
#pragma omp parallel for schedule(static) reduction(+:Fx,Fy,Fz)
    for (int x = 0; x < X; ++x)
        for (int y = 0; y < Y; ++y)
            for (int z = 0; z < Z; ++z) {
                const size_t k = lbo_index(x, y, z, Y, Z);
                float f[27];
                if (solid && solid[k]) {
                    for (int i = 0; i < q; ++i) fout[(size_t)i * n + k] = 0.0f;
                    continue;
                }
                for (int i = 0; i < q; ++i) {
                    int s[3] = {x - LBO_DIRECTIONS[i][0], y - LBO_DIRECTIONS[i][1], z - LBO_DIRECTIONS[i][2]};
                    int wall_axis = -1, wall_side = 0, zero = 0;
                    for (int a = 0; a < 3; ++a) {
                        if (s[a] >= 0 && s[a] < dims[a]) continue;
                        if (bc[a] == 2) { if (wall_axis < 0) { wall_axis = a; wall_side = (s[a] < 0) ? 0 : 1; } }
                        else if (bc[a] == 0) zero = 1;
                        else s[a] = (s[a] + dims[a]) % dims[a];
                    }
                    if (wall_axis >= 0) {
                        /* half-way bounce-back off a (possibly moving) domain wall, rho_wall = 1 */
                        const float *uw = wall_u ? wall_u[wall_axis * 2 + wall_side] : NULL;
                        float cu = uw ? (LBO_DIRECTIONS[i][0] * uw[0] + LBO_DIRECTIONS[i][1] * uw[1] + LBO_DIRECTIONS[i][2] * uw[2]) : 0.0f;
                        f[i] = fin[(size_t)LBO_INVERSE[i] * n + k] + 6.0f * (float)q_weight(q, i) * cu;
                    } else if (zero) {
                        f[i] = 0.0f;
                    } else {
                        size_t ks = lbo_index(s[0], s[1], s[2], Y, Z);
                        if (solid && solid[ks]) {
                            /* bounce-back off an obstacle cell; the momentum it takes is -c_i (f_in + f_out) */
                            float back = fin[(size_t)LBO_INVERSE[i] * n + k];
                            f[i] = back;
                            Fx -= LBO_DIRECTIONS[i][0] * (double)(back + back);
                            Fy -= LBO_DIRECTIONS[i][1] * (double)(back + back);
                            Fz -= LBO_DIRECTIONS[i][2] * (double)(back + back);
                        } else {
                            f[i] = fin[(size_t)i * n + ks];
                        }
                    }
                }
                collide_cell_ext(f, q, tau, moment, eq64);
                for (int i = 0; i < q; ++i) fout[(size_t)i * n + k] = f[i];
            }
    if (force) { force[0] += Fx; force[1] += Fy; force[2] += Fz; }
}

// lbm_geometry.cu -- the side inputs of the hot path, produced on the GPU: the solid mask
// (Lattice::insertModel, simulation/source/Lattice.cpp:228-260) and the reflection tables
// (Lattice::preProcessModel, :323-426), plus the small geometry helpers the reference's API exposes
// (Model, simulation/source/Model.cpp; rotateVectorAroundAxis, Geometry.cpp:22-42; distributeVector,
// Lattice.cpp:262-321; compressInts, Utils.cpp:4-29).
//
// The reference asks CGAL for "is this point inside the mesh" and "which facet does this ray hit first"
// (Model.cpp:51-63,75-91), one query at a time on one CPU thread: 1.3e8 queries for a 512^3 lattice.
// Here the same questions are answered by ray casting in double arithmetic on float inputs, one thread per
// lattice cell / per boundary cell.  A point lying on a facet counts as inside (ON_BOUNDARY, pinned by
// test/CoordinatesTest.cpp:137-169); a ray that grazes an edge or vertex is re-shot in another direction.
// This file is compiled with -fmad=false so that host and device evaluate every predicate identically and
// the mask and the tables are reproducible bit for bit.
#include "lbm_internal.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace lbm;

struct lbm_mesh_s {
    std::vector<float> v;          // 3 per vertex
    std::vector<int> face_start;   // faces + 1
    std::vector<int> face_idx;
    std::vector<int> tri;          // 3 per triangular face, in face order
    bool loaded = false, valid = true, closed = true;
    float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    // device copies
    float* d_v = nullptr;
    int* d_tri = nullptr;
    int device = -1;
};

namespace {

#define GCK(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char b_[256];                                                                          \
            snprintf(b_, sizeof b_, "%s: %s", #call, cudaGetErrorString(e_));                      \
            return lbm_internal_fail(LBM_ERR_CUDA, b_);                                            \
        }                                                                                          \
    } while (0)

struct MeshView {
    const float* v;
    const int* tri;
    int nf;
    int valid;
    float lo[3], hi[3];
};

// ---- predicates (host and device) --------------------------------------------------------------------

LBM_HD void load_tri(const MeshView& m, int f, double a[3], double b[3], double c[3])
{
    const int* id = m.tri + 3 * f;
    for (int k = 0; k < 3; ++k) {
        a[k] = m.v[id[0] * 3 + k];
        b[k] = m.v[id[1] * 3 + k];
        c[k] = m.v[id[2] * 3 + k];
    }
}

LBM_HD void cross(const double a[3], const double b[3], double o[3])
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

LBM_HD double dot(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// p on the closed triangle f
LBM_HD bool on_triangle(const MeshView& m, int f, const double p[3])
{
    double a[3], b[3], c[3], e1[3], e2[3], n[3], d[3];
    load_tri(m, f, a, b, c);
    for (int k = 0; k < 3; ++k) { e1[k] = b[k] - a[k]; e2[k] = c[k] - a[k]; d[k] = p[k] - a[k]; }
    cross(e1, e2, n);
    const double nn = dot(n, n);
    if (nn == 0.0) return false;
    const double dist = dot(n, d);
    const double scale = sqrt(nn) * (sqrt(dot(d, d)) + 1e-30);
    if (fabs(dist) > 1e-12 * scale) return false;
    const double d11 = dot(e1, e1), d12 = dot(e1, e2), d22 = dot(e2, e2), p1 = dot(d, e1), p2 = dot(d, e2);
    const double det = d11 * d22 - d12 * d12;
    const double u = (d22 * p1 - d12 * p2) / det, v = (d11 * p2 - d12 * p1) / det;
    const double eps = 1e-9;
    return u >= -eps && v >= -eps && u + v <= 1.0 + eps;
}

// 0 miss, 1 clean hit, 2 degenerate (in-plane, or through an edge / vertex / the origin)
LBM_HD int ray_triangle(const MeshView& m, int f, const double o[3], const double dir[3], double* t_out)
{
    double a[3], b[3], c[3], e1[3], e2[3], pv[3], tv[3], qv[3];
    load_tri(m, f, a, b, c);
    for (int k = 0; k < 3; ++k) { e1[k] = b[k] - a[k]; e2[k] = c[k] - a[k]; tv[k] = o[k] - a[k]; }
    cross(dir, e2, pv);
    const double det = dot(e1, pv);
    const double scale = sqrt(dot(e1, e1) * dot(e2, e2) * dot(dir, dir)) + 1e-300;
    const double eps = 1e-9;
    if (fabs(det) <= 1e-12 * scale) {
        double n[3];
        cross(e1, e2, n);
        const double off = dot(n, tv);
        if (fabs(off) <= 1e-12 * sqrt(dot(n, n)) * (sqrt(dot(tv, tv)) + 1e-30)) return 2;
        return 0;
    }
    const double inv = 1.0 / det;
    const double u = dot(tv, pv) * inv;
    cross(tv, e1, qv);
    const double v = dot(dir, qv) * inv;
    const double t = dot(e2, qv) * inv;
    if (u < -eps || v < -eps || u + v > 1.0 + eps) return 0;
    if (t < -eps) return 0;
    *t_out = t;
    if (u < eps || v < eps || u + v > 1.0 - eps || t < eps) return 2;
    return 1;
}

// Model::isPointInsideModel: bounded side or boundary
LBM_HD bool point_inside(const MeshView& m, float px, float py, float pz)
{
    if (!m.valid || m.nf == 0) return false;
    if (px < m.lo[0] || px > m.hi[0] || py < m.lo[1] || py > m.hi[1] || pz < m.lo[2] || pz > m.hi[2]) return false;
    const double p[3] = {px, py, pz};
    for (int f = 0; f < m.nf; ++f)
        if (on_triangle(m, f, p)) return true;
    unsigned long long s = 0x9E3779B97F4A7C15ull;
    for (int attempt = 0; attempt < 64; ++attempt) {
        double dir[3];
        if (attempt == 0) { dir[0] = 0.0; dir[1] = 0.0; dir[2] = 1.0; }
        else {
            for (int k = 0; k < 3; ++k) {
                s = s * 6364136223846793005ull + 1442695040888963407ull;
                dir[k] = ((double)(s >> 11) / 9007199254740992.0) * 2.0 - 1.0;
            }
        }
        int crossings = 0;
        bool bad = false;
        for (int f = 0; f < m.nf && !bad; ++f) {
            double t;
            const int r = ray_triangle(m, f, p, dir, &t);
            if (r == 2) bad = true;
            else if (r == 1) ++crossings;
        }
        if (!bad) return crossings & 1;
    }
    return false;
}

// The same answer from far fewer triangles.  A triangle can only matter to the point (px, py, pz) if
//   * its xy bounding box (slightly inflated) holds (px, py): only then can the point lie on it or the +z ray hit it, or
//   * it is parallel to the +z ray ("vertical"): ray_triangle then reports a degenerate hit for EVERY point in its plane,
//     however far from the triangle, and the ray is re-shot.
// `cand` = the triangles binned to this lattice column, `vert` = all vertical triangles.  Every triangle left out would
// have answered "no" to both tests, so the crossings counted are the ones point_inside() counts; a degenerate hit falls
// back to point_inside() itself (re-shot rays go anywhere and need every triangle).
LBM_HD bool point_inside_binned(const MeshView& m, float px, float py, float pz, const int* cand, int ncand, const int* vert, int nvert)
{
    if (!m.valid || m.nf == 0) return false;
    if (px < m.lo[0] || px > m.hi[0] || py < m.lo[1] || py > m.hi[1] || pz < m.lo[2] || pz > m.hi[2]) return false;
    const double p[3] = {px, py, pz};
    for (int k = 0; k < ncand; ++k)
        if (on_triangle(m, cand[k], p)) return true;
    const double dir[3] = {0.0, 0.0, 1.0};
    int crossings = 0;
    bool bad = false;
    for (int k = 0; k < ncand && !bad; ++k) {
        double t;
        const int r = ray_triangle(m, cand[k], p, dir, &t);
        if (r == 2) bad = true;
        else if (r == 1) ++crossings;
    }
    for (int k = 0; k < nvert && !bad; ++k) {
        double t;
        if (ray_triangle(m, vert[k], p, dir, &t) == 2) bad = true;
    }
    if (!bad) return crossings & 1;
    return point_inside(m, px, py, pz);
}

// is triangle f parallel to the +z ray in ray_triangle's own terms (its first branch)?
LBM_HD bool parallel_to_z(const MeshView& m, int f)
{
    double a[3], b[3], c[3], e1[3], e2[3], pv[3];
    load_tri(m, f, a, b, c);
    for (int k = 0; k < 3; ++k) { e1[k] = b[k] - a[k]; e2[k] = c[k] - a[k]; }
    const double dir[3] = {0.0, 0.0, 1.0};
    cross(dir, e2, pv);
    const double det = dot(e1, pv);
    const double scale = sqrt(dot(e1, e1) * dot(e2, e2) * dot(dir, dir)) + 1e-300;
    return fabs(det) <= 1e-12 * scale;
}

// rotateVectorAroundAxis with the half-angle cosine / sine supplied (they come from the host's libm so
// that host and device agree to the last bit)
LBM_HD void rotate_cs(const float axis_in[3], const float vec[3], float c, float s, float cn, float sn, float out[3])
{
    float axis[3] = {axis_in[0], axis_in[1], axis_in[2]};
    const float magnitude = sqrtf((axis[0] * axis[0]) + (axis[1] * axis[1]) + (axis[2] * axis[2]));
    axis[0] /= magnitude; axis[1] /= magnitude; axis[2] /= magnitude;
    const float q[4] = {c, s * axis[0], s * axis[1], s * axis[2]};
    const float qi[4] = {cn, sn * axis[0], sn * axis[1], sn * axis[2]};
    const float pt[4] = {0.0f, vec[0], vec[1], vec[2]};
    float t[4], r[4];
    // quaternion products q * pt * q^-1
    t[0] = (q[0] * pt[0]) - (q[1] * pt[1]) - (q[2] * pt[2]) - (q[3] * pt[3]);
    t[1] = (q[0] * pt[1]) + (q[1] * pt[0]) + (q[2] * pt[3]) - (q[3] * pt[2]);
    t[2] = (q[0] * pt[2]) - (q[1] * pt[3]) + (q[2] * pt[0]) + (q[3] * pt[1]);
    t[3] = (q[0] * pt[3]) + (q[1] * pt[2]) - (q[2] * pt[1]) + (q[3] * pt[0]);
    r[1] = (t[0] * qi[1]) + (t[1] * qi[0]) + (t[2] * qi[3]) - (t[3] * qi[2]);
    r[2] = (t[0] * qi[2]) - (t[1] * qi[3]) + (t[2] * qi[0]) + (t[3] * qi[1]);
    r[3] = (t[0] * qi[3]) + (t[1] * qi[2]) - (t[2] * qi[1]) + (t[3] * qi[0]);
    out[0] = r[1]; out[1] = r[2]; out[2] = r[3];
}

struct HalfTurn { float c, s, cn, sn; };   // cos / sin of +pi/2 and -pi/2 as the host's libm rounds them

// one triangle of the first-hit search of Model::reflectionVector (AABB_tree::first_intersection in the reference)
LBM_HD void first_hit_consider(const MeshView& m, int f, const double o[3], const double dir[3], int& best, double& best_t)
{
    double t;
    const int r = ray_triangle(m, f, o, dir, &t);
    if (r == 0) return;
    if (r == 2) {
        double a[3], b[3], c[3], e1[3], e2[3], n[3];
        load_tri(m, f, a, b, c);
        for (int k = 0; k < 3; ++k) { e1[k] = b[k] - a[k]; e2[k] = c[k] - a[k]; }
        cross(e1, e2, n);
        if (fabs(dot(n, dir)) <= 1e-12 * sqrt(dot(n, n) * dot(dir, dir))) return;   // in-plane ray: no usable hit
    }
    if (best < 0 || t < best_t - 1e-12) { best = f; best_t = t; }
}

// the mirror of the incoming direction p1 -> p2 about facet `best` (none: the zero vector)
LBM_HD void mirror_about_facet(const MeshView& m, int best, const float p1[3], const float p2[3], HalfTurn ht, float out[3])
{
    out[0] = out[1] = out[2] = 0.0f;
    float vx = p1[0] - p2[0], vy = p1[1] - p2[1], vz = p1[2] - p2[2];
    const float norm = sqrtf((vx * vx) + (vy * vy) + (vz * vz));
    vx /= norm; vy /= norm; vz /= norm;
    if (best < 0) return;
    const int* id = m.tri + 3 * best;
    const float* A = m.v + id[0] * 3;
    const float* B = m.v + id[1] * 3;
    const float* C = m.v + id[2] * 3;
    const float e1[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
    const float e2[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
    float n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    const float nl = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    if (nl != 0.0f) { n[0] /= nl; n[1] /= nl; n[2] /= nl; }
    const float v[3] = {vx, vy, vz};
    rotate_cs(n, v, ht.c, ht.s, ht.cn, ht.sn, out);
}

// Model::reflectionVector: mirror of the incoming direction about the first facet the ray p1 -> p2 hits
LBM_HD void reflection_vector(const MeshView& m, const float p1[3], const float p2[3], HalfTurn ht, float out[3])
{
    out[0] = out[1] = out[2] = 0.0f;
    if (m.nf == 0) return;
    const double o[3] = {p1[0], p1[1], p1[2]};
    const double dir[3] = {(double)p2[0] - p1[0], (double)p2[1] - p1[1], (double)p2[2] - p1[2]};
    int best = -1;
    double best_t = 0.0;
    for (int f = 0; f < m.nf; ++f) first_hit_consider(m, f, o, dir, best, best_t);
    mirror_about_facet(m, best, p1, p2, ht, out);
}

// The same for a ray between two NEIGHBOURING lattice points with the triangles binned per lattice column.  A triangle
// that the ray meets between the two points (t <= 1) covers, in xy, a point of the segment between two adjacent columns,
// and every column's bin reaches one column further than the triangle's box: it is in the bin of p1's column or of p2's.
// So the two bins are searched first, in ascending triangle order like the full loop (ties go to the lower index); a
// hit at t <= 1 found there is the first hit of the whole mesh.  Only a ray that meets nothing between the two points
// (it then runs on through other columns) takes the full loop.
LBM_HD void reflection_vector_binned(const MeshView& m, const float p1[3], const float p2[3], HalfTurn ht, float out[3],
                                     const int* bin1, int n1, const int* bin2, int n2)
{
    out[0] = out[1] = out[2] = 0.0f;
    if (m.nf == 0) return;
    const double o[3] = {p1[0], p1[1], p1[2]};
    const double dir[3] = {(double)p2[0] - p1[0], (double)p2[1] - p1[1], (double)p2[2] - p1[2]};
    int best = -1;
    double best_t = 0.0;
    int i = 0, j = 0;
    while (i < n1 || j < n2) {                 // merge of two ascending lists, every triangle once
        int f;
        if (j >= n2 || (i < n1 && bin1[i] <= bin2[j])) { f = bin1[i]; if (j < n2 && bin2[j] == f) ++j; ++i; }
        else { f = bin2[j]; ++j; }
        first_hit_consider(m, f, o, dir, best, best_t);
    }
    if (best < 0 || best_t > 1.0 - 1e-9) {     // nothing (safely) between the two points: ask the whole mesh
        best = -1;
        best_t = 0.0;
        for (int f = 0; f < m.nf; ++f) first_hit_consider(m, f, o, dir, best, best_t);
    }
    mirror_about_facet(m, best, p1, p2, ht, out);
}

// Lattice::distributeVector: the 27 unit directions ranked by 2 - distance, ties broken towards the
// larger index (a max-heap of (score, index) pairs popped 27 times)
LBM_HD void distribute_vector(const float v[3], float scores[27], int dirs[27])
{
    float sc[27];
    for (int i = 0; i < 27; ++i) {
        const int nz = (dir_c(i, 0) != 0) + (dir_c(i, 1) != 0) + (dir_c(i, 2) != 0);
        const double len = (nz == 2) ? 1.4142135623730951 : (nz == 3) ? 1.7320508075688772 : 1.0;
        float d2 = 0.0f;
        float comp[3];
        for (int a = 0; a < 3; ++a) {
            const float unit = (dir_c(i, a) == 0) ? 0.0f : (float)((double)dir_c(i, a) / len);
            comp[a] = v[a] - unit;
        }
        d2 = (comp[0] * comp[0]) + (comp[1] * comp[1]) + (comp[2] * comp[2]);
        sc[i] = (float)(2.0 - sqrtf(d2));
    }
    unsigned used = 0;
    for (int r = 0; r < 27; ++r) {
        int best = -1;
        for (int i = 0; i < 27; ++i) {
            if (used & (1u << i)) continue;
            if (best < 0 || sc[i] > sc[best] || (sc[i] == sc[best] && i > best)) best = i;
        }
        used |= 1u << best;
        scores[r] = sc[best];
        dirs[r] = best;
    }
}

// ---- kernels ----------------------------------------------------------------------------------------------

// Lattice::insertModel's triple loop (Lattice.cpp:248-257), one thread per cell, host cell order
// (slabs: plane 0 of the array is global plane x_origin of gx; planes outside the lattice are fluid)
__global__ void voxelise_kernel(MeshView m, uint8_t* __restrict__ solid, int nx, int ny, int nz, float spacing, int x_origin, int gx)
{
    const long long n = (long long)nx * ny * nz;
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int z = (int)(k % nz);
    const int y = (int)((k / nz) % ny);
    const int x = (int)(k / ((long long)nz * ny)) + x_origin;
    solid[k] = (x >= 0 && x < gx && point_inside(m, x * spacing, y * spacing, z * spacing)) ? 1 : 0;
}

// The same loop with the triangles binned per lattice column (x, y): col_start / col_tris is a CSR list over the nx * ny
// columns of the planes being voxelised; cost per cell = the triangles over its column instead of the whole mesh.
__global__ void voxelise_binned_kernel(MeshView m, const int* __restrict__ col_start, const int* __restrict__ col_tris,
                                       const int* __restrict__ vert, int nvert, uint8_t* __restrict__ solid,
                                       int nx, int ny, int nz, float spacing, int x_origin, int gx)
{
    const long long n = (long long)nx * ny * nz;
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int z = (int)(k % nz);
    const int y = (int)((k / nz) % ny);
    const int xl = (int)(k / ((long long)nz * ny));
    const int x = xl + x_origin;
    const int c0 = col_start[xl * ny + y], c1 = col_start[xl * ny + y + 1];
    solid[k] = (x >= 0 && x < gx && point_inside_binned(m, x * spacing, y * spacing, z * spacing, col_tris + c0, c1 - c0, vert, nvert)) ? 1 : 0;
}

// solid cells with at least one in-domain fluid neighbour: the only cells that can own a table
__global__ void candidate_kernel(const uint8_t* __restrict__ solid, uint8_t* __restrict__ cand, int nx, int ny, int nz)
{
    const long long n = (long long)nx * ny * nz;
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint8_t c = 0;
    if (solid[k]) {
        const int z = (int)(k % nz), y = (int)((k / nz) % ny), x = (int)(k / ((long long)nz * ny));
        for (int d = 1; d < 27 && !c; ++d) {
            const int tx = x + dir_c(d, 0), ty = y + dir_c(d, 1), tz = z + dir_c(d, 2);
            if (tx < 0 || ty < 0 || tz < 0 || tx >= nx || ty >= ny || tz >= nz) continue;
            if (!solid[((long long)tx * ny + ty) * nz + tz]) c = 1;
        }
    }
    cand[k] = c;
}

struct TableRecord {       // ReflectionValues, simulation/include/LatticePoint.h:3-7
    int32_t dirs[14];
    float w[81];
};

// Lattice::preProcessModel's per-cell body (Lattice.cpp:342-411), one thread per candidate cell.
// An out-of-domain neighbour is treated as unavailable where the reference indexes without a check (:369).
// (slabs: the array holds the planes x_origin .. x_origin + nx - 1 of the lattice; positions use the global index)
__global__ void table_kernel(MeshView m, const uint8_t* __restrict__ solid, const long long* __restrict__ cells, int ncells,
                             TableRecord* __restrict__ tables, uint8_t* __restrict__ reflected, int nx, int ny, int nz,
                             float spacing, HalfTurn ht, int x_origin, const int* __restrict__ col_start, const int* __restrict__ col_tris)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const long long k = cells[c];
    const int z = (int)(k % nz), y = (int)((k / nz) % ny), x = (int)(k / ((long long)nz * ny));
    TableRecord& t = tables[c];
    for (int i = 0; i < 14; ++i) t.dirs[i] = 0;
    for (int i = 0; i < 81; ++i) t.w[i] = 0.0f;
    bool any = false;
    for (int d = 0; d < 27; ++d) {
        const int nxp = x + dir_c(d, 0), nyp = y + dir_c(d, 1), nzp = z + dir_c(d, 2);
        if (nxp < 0 || nyp < 0 || nzp < 0 || nxp >= nx || nyp >= ny || nzp >= nz) continue;
        if (solid[((long long)nxp * ny + nyp) * nz + nzp]) continue;
        const float p1[3] = {spacing * (nxp + x_origin), spacing * nyp, spacing * nzp};
        const float p2[3] = {spacing * (x + x_origin), spacing * y, spacing * z};
        float r[3], scores[27];
        int dirs[27];
        if (col_start) {
            const int c1 = nxp * ny + nyp, c2 = x * ny + y;          // columns of the fluid neighbour and of this cell
            reflection_vector_binned(m, p1, p2, ht, r, col_tris + col_start[c1], col_start[c1 + 1] - col_start[c1],
                                     col_tris + col_start[c2], col_start[c2 + 1] - col_start[c2]);
        } else {
            reflection_vector(m, p1, p2, ht, r);
        }
        distribute_vector(r, scores, dirs);
        int counter = 0;
        float norm = 0.0f;
        const int base = dir_inv(d) * 3;
        for (int i = 0; i < 27 && counter < 3; ++i) {
            const int cd = dirs[i];
            const int tx = x + dir_c(cd, 0), ty = y + dir_c(cd, 1), tz = z + dir_c(cd, 2);
            if (tx < 0 || ty < 0 || tz < 0 || tx >= nx || ty >= ny || tz >= nz) continue;
            if (solid[((long long)tx * ny + ty) * nz + tz]) continue;
            const int slot = base + counter;
            // 5-bit packing, 6 slots per int (Lattice.cpp:450-463)
            const unsigned shift = 5u * (unsigned)(slot % 6);
            unsigned word = (unsigned)t.dirs[slot / 6];
            word = (word & ~(0x1Fu << shift)) | ((unsigned)cd << shift);
            t.dirs[slot / 6] = (int32_t)word;
            t.w[slot] = scores[i];
            norm += scores[i];
            any = true;
            ++counter;
        }
        if (norm != 0.0f) {
            for (int i = 0; i < 3; ++i) {
                const int slot = base + i;
                const unsigned dd = ((unsigned)t.dirs[slot / 6] >> (5u * (unsigned)(slot % 6))) & 0x1Fu;
                if (dd != 0) t.w[slot] /= norm;
            }
        }
    }
    reflected[c] = any ? 1 : 0;
}

MeshView host_view(const lbm_mesh_s* m)
{
    MeshView v;
    v.v = m->v.data();
    v.tri = m->tri.data();
    v.nf = (int)(m->tri.size() / 3);
    v.valid = (m->valid && m->loaded) ? 1 : 0;
    for (int a = 0; a < 3; ++a) { v.lo[a] = m->lo[a]; v.hi[a] = m->hi[a]; }
    return v;
}

int device_view(lbm_mesh_s* m, int device, MeshView* out)
{
    if (m->device != device) {
        cudaFree(m->d_v);
        cudaFree(m->d_tri);
        m->d_v = nullptr;
        m->d_tri = nullptr;
        const size_t vb = std::max<size_t>(m->v.size(), 1) * sizeof(float), tb = std::max<size_t>(m->tri.size(), 1) * sizeof(int);
        GCK(cudaMalloc(&m->d_v, vb));
        GCK(cudaMalloc(&m->d_tri, tb));
        if (!m->v.empty()) GCK(cudaMemcpy(m->d_v, m->v.data(), m->v.size() * sizeof(float), cudaMemcpyHostToDevice));
        if (!m->tri.empty()) GCK(cudaMemcpy(m->d_tri, m->tri.data(), m->tri.size() * sizeof(int), cudaMemcpyHostToDevice));
        m->device = device;
    }
    *out = host_view(m);
    out->v = m->d_v;
    out->tri = m->d_tri;
    return LBM_OK;
}

HalfTurn half_turn()
{
    // Model.cpp:86 passes M_PI as a float angle; Geometry.cpp:32 halves it in double and narrows again
    float angle = (float)M_PI;
    angle = (float)(angle / 2.0);
    HalfTurn h;
    h.c = cosf(angle); h.s = sinf(angle); h.cn = cosf(-angle); h.sn = sinf(-angle);
    return h;
}

}  // namespace

// ---- C ABI ---------------------------------------------------------------------------------------------------

extern "C" {

// Model::importModel (Model.cpp:22-44): OBJ vertices and faces; valid = all faces triangles and the surface closed
LBM_API int lbm_mesh_load(const char* path, lbm_mesh* out)
{
    if (!path || !out) return lbm_internal_fail(LBM_ERR_INVALID, "null argument");
    lbm_mesh_s* m = new lbm_mesh_s();
    m->face_start.push_back(0);
    *out = m;
    FILE* fp = fopen(path, "r");
    if (!fp) return LBM_OK;          // the reference returns silently and keeps an empty model (Model.cpp:24-27)
    m->loaded = true;
    char line[4096];
    while (fgets(line, sizeof line, fp)) {
        if (char* hash = strchr(line, '#')) *hash = 0;
        char* p = line;
        while (isspace((unsigned char)*p)) ++p;
        if (p[0] == 'v' && isspace((unsigned char)p[1])) {
            double a, b, c;
            if (sscanf(p + 1, "%lf %lf %lf", &a, &b, &c) == 3) {
                m->v.push_back((float)a); m->v.push_back((float)b); m->v.push_back((float)c);
            }
        } else if (p[0] == 'f' && isspace((unsigned char)p[1])) {
            ++p;
            std::vector<int> ids;
            for (;;) {
                while (isspace((unsigned char)*p)) ++p;
                if (!*p) break;
                char* end;
                const long id = strtol(p, &end, 10);
                if (end == p) break;
                while (*end && !isspace((unsigned char)*end)) ++end;      // skip /vt/vn
                p = end;
                ids.push_back(id > 0 ? (int)id - 1 : (int)(m->v.size() / 3) + (int)id);
            }
            if (ids.size() >= 3) {
                m->face_idx.insert(m->face_idx.end(), ids.begin(), ids.end());
                m->face_start.push_back((int)m->face_idx.size());
            }
        }
    }
    fclose(fp);
    const int nf = (int)m->face_start.size() - 1;
    bool triangles = true;
    // a face that names a vertex the file does not have (`f 0 ...`, an index beyond the vertex list) makes the mesh
    // invalid -- nothing is inside, like any mesh CGAL's reader rejects -- instead of an out-of-range read later
    bool indices_ok = true;
    const int nv_all = (int)(m->v.size() / 3);
    for (int id : m->face_idx)
        if (id < 0 || id >= nv_all) indices_ok = false;
    if (!indices_ok) {
        m->face_idx.clear();
        m->face_start.assign(1, 0);
        m->closed = false;
        m->valid = false;
        return LBM_OK;
    }
    std::map<std::pair<int, int>, int> edges;
    for (int f = 0; f < nf; ++f) {
        const int b = m->face_start[f], e = m->face_start[f + 1];
        if (e - b != 3) triangles = false;
        else m->tri.insert(m->tri.end(), m->face_idx.begin() + b, m->face_idx.begin() + e);
        for (int j = b; j < e; ++j) {
            const int a0 = m->face_idx[j], a1 = m->face_idx[j + 1 < e ? j + 1 : b];
            edges[{std::min(a0, a1), std::max(a0, a1)}] += 1;
        }
    }
    m->closed = true;
    for (const auto& kv : edges)
        if (kv.second != 2) m->closed = false;
    m->valid = triangles && m->closed;                    // Model.cpp:31,40-43
    const int nv = (int)(m->v.size() / 3);
    for (int i = 0; i < nv; ++i)
        for (int a = 0; a < 3; ++a) {
            const float c = m->v[i * 3 + a];
            if (i == 0 || c < m->lo[a]) m->lo[a] = c;
            if (i == 0 || c > m->hi[a]) m->hi[a] = c;
        }
    return LBM_OK;
}

LBM_API int lbm_mesh_destroy(lbm_mesh m)
{
    if (!m) return LBM_OK;
    cudaFree(m->d_v);
    cudaFree(m->d_tri);
    delete m;
    return LBM_OK;
}

LBM_API int lbm_mesh_info(lbm_mesh m, int* n_vertices, int* n_faces, int* closed, int* valid)
{
    if (!m) return lbm_internal_fail(LBM_ERR_INVALID, "null mesh");
    if (n_vertices) *n_vertices = (int)(m->v.size() / 3);
    if (n_faces) *n_faces = (int)m->face_start.size() - 1;
    if (closed) *closed = m->closed ? 1 : 0;
    if (valid) *valid = m->valid ? 1 : 0;
    return LBM_OK;
}

LBM_API int lbm_mesh_bounding_box(lbm_mesh m, float lo[3], float hi[3])
{
    if (!m || !lo || !hi) return lbm_internal_fail(LBM_ERR_INVALID, "null argument");
    for (int a = 0; a < 3; ++a) { lo[a] = m->lo[a]; hi[a] = m->hi[a]; }
    return LBM_OK;
}

LBM_API int lbm_mesh_point_inside(lbm_mesh m, const float p[3], int* inside)
{
    if (!m || !p || !inside) return lbm_internal_fail(LBM_ERR_INVALID, "null argument");
    *inside = point_inside(host_view(m), p[0], p[1], p[2]) ? 1 : 0;
    return LBM_OK;
}

LBM_API int lbm_mesh_reflection_vector(lbm_mesh m, const float p1[3], const float p2[3], float out[3])
{
    if (!m || !p1 || !p2 || !out) return lbm_internal_fail(LBM_ERR_INVALID, "null argument");
    MeshView v = host_view(m);
    if (!m->loaded) v.nf = 0;
    reflection_vector(v, p1, p2, half_turn(), out);
    return LBM_OK;
}

LBM_API int lbm_rotate_vector_around_axis(const float axis[3], const float vec[3], float angle, float out[3])
{
    if (!axis || !vec || !out) return lbm_internal_fail(LBM_ERR_INVALID, "null argument");
    angle = (float)(angle / 2.0);
    rotate_cs(axis, vec, cosf(angle), sinf(angle), cosf(-angle), sinf(-angle), out);
    return LBM_OK;
}

LBM_API int lbm_distribute_vector(const float v[3], float scores[27], int32_t dirs[27])
{
    if (!v || !scores || !dirs) return lbm_internal_fail(LBM_ERR_INVALID, "null argument");
    int d[27];
    distribute_vector(v, scores, d);
    for (int i = 0; i < 27; ++i) dirs[i] = d[i];
    return LBM_OK;
}

// Utils::compressInts (Utils.cpp:4-29): first element in the highest bits, the shift is always 5
LBM_API unsigned lbm_compress_ints(const int* array, int num, int bit_length)
{
    const unsigned limit = 1u << bit_length;
    unsigned out = 0;
    for (int i = 0; i < num; ++i)
        if ((unsigned)array[i] < limit) {
            out |= (unsigned)array[i];
            if (i < num - 1) out <<= 5;
        }
    return out;
}

// The triangles of a mesh binned per lattice column (x, y) of the planes x_origin .. x_origin + nxv - 1 (global indices):
// each triangle goes to the columns its inflated xy bounding box covers plus one column all round (rounding of
// index * spacing; the table builder relies on that extra column too).  CSR on the device + the list of triangles
// parallel to the z axis.  `ok` false: no bins (empty or invalid mesh, switched off, or they would not fit).
struct DeviceBins {
    int *start = nullptr, *tris = nullptr, *vert = nullptr;
    int nvert = 0;
    bool ok = false;
    void release() { cudaFree(start); cudaFree(tris); cudaFree(vert); start = tris = vert = nullptr; ok = false; }
};

static int build_bins(lbm_handle h, lbm_mesh m, int nxv, int x_origin, float spacing, DeviceBins* out)
{
    const int ny = h->g.ny;
    static const bool binned = [] { const char* e = getenv("LBM_VOXELISE_BINNED"); return e ? atoi(e) != 0 : true; }();
    const int nf = (int)(m->tri.size() / 3);
    const size_t ncol = (size_t)nxv * ny;
    if (!binned || nf == 0 || !m->valid || !m->loaded) return LBM_OK;
    std::vector<int> count(ncol + 1, 0), vertical;
    std::vector<int> lo_x(nf), hi_x(nf), lo_y(nf), hi_y(nf);
    size_t entries = 0;
    const MeshView hv = host_view(m);
    for (int f = 0; f < nf; ++f) {
        if (parallel_to_z(hv, f)) vertical.push_back(f);
        double bl[3] = {1e300, 1e300, 1e300}, bh[3] = {-1e300, -1e300, -1e300};
        for (int c = 0; c < 3; ++c)
            for (int a = 0; a < 3; ++a) {
                const double w = m->v[(size_t)m->tri[3 * f + c] * 3 + a];
                bl[a] = std::min(bl[a], w); bh[a] = std::max(bh[a], w);
            }
        const double diag = sqrt((bh[0] - bl[0]) * (bh[0] - bl[0]) + (bh[1] - bl[1]) * (bh[1] - bl[1]) + (bh[2] - bl[2]) * (bh[2] - bl[2]));
        const double margin = 1e-6 * (diag + fabs(bl[0]) + fabs(bh[0]) + fabs(bl[1]) + fabs(bh[1])) + 1e-30;
        // columns whose lattice point can lie within the inflated box, one more on each side
        const long long ax = (long long)floor((bl[0] - margin) / spacing) - 1 - x_origin, bx = (long long)ceil((bh[0] + margin) / spacing) + 1 - x_origin;
        const long long ay = (long long)floor((bl[1] - margin) / spacing) - 1, by = (long long)ceil((bh[1] + margin) / spacing) + 1;
        lo_x[f] = (int)std::max<long long>(ax, 0); hi_x[f] = (int)std::min<long long>(bx, nxv - 1);
        lo_y[f] = (int)std::max<long long>(ay, 0); hi_y[f] = (int)std::min<long long>(by, ny - 1);
        if (hi_x[f] < lo_x[f] || hi_y[f] < lo_y[f]) continue;
        entries += (size_t)(hi_x[f] - lo_x[f] + 1) * (size_t)(hi_y[f] - lo_y[f] + 1);
    }
    if (entries > (size_t(64) << 20)) return LBM_OK;          // huge triangles on a huge lattice: all-pairs kernels
    for (int f = 0; f < nf; ++f)
        for (int x = lo_x[f]; x <= hi_x[f]; ++x)
            for (int y = lo_y[f]; y <= hi_y[f]; ++y) ++count[(size_t)x * ny + y + 1];
    for (size_t c = 0; c < ncol; ++c) count[c + 1] += count[c];
    std::vector<int> tris(entries ? entries : 1), fill(count.begin(), count.end() - 1);
    for (int f = 0; f < nf; ++f)             // ascending triangle order within a column, like the all-pairs loops
        for (int x = lo_x[f]; x <= hi_x[f]; ++x)
            for (int y = lo_y[f]; y <= hi_y[f]; ++y) tris[(size_t)fill[(size_t)x * ny + y]++] = f;
    cudaError_t e = cudaMalloc(&out->start, count.size() * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&out->tris, tris.size() * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&out->vert, (vertical.size() + 1) * sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpyAsync(out->start, count.data(), count.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out->tris, tris.data(), tris.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess && !vertical.empty())
        e = cudaMemcpyAsync(out->vert, vertical.data(), vertical.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);      // the host vectors go away here
    if (e != cudaSuccess) {
        out->release();
        return lbm_internal_fail(e == cudaErrorMemoryAllocation ? LBM_ERR_NOMEM : LBM_ERR_CUDA, cudaGetErrorString(e));
    }
    out->nvert = (int)vertical.size();
    out->ok = true;
    return LBM_OK;
}

// Voxelises the planes x_origin .. x_origin + nxv - 1 (global indices) into `bytes` ([nxv][ny][nz], device): with the bins a
// cell looks at the triangles over its column only, without them at the whole mesh.  `keep` (optional) receives the bins.
static int voxelise_planes(lbm_handle h, lbm_mesh m, const MeshView& v, uint8_t* bytes, int nxv, int x_origin, int gx, float spacing,
                           DeviceBins* keep = nullptr)
{
    const Geom& g = h->g;
    const size_t n = (size_t)nxv * g.ny * g.nz;
    DeviceBins bins;
    int rc = build_bins(h, m, nxv, x_origin, spacing, &bins);
    if (rc) return rc;
    if (bins.ok)
        voxelise_binned_kernel<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(v, bins.start, bins.tris, bins.vert, bins.nvert, bytes,
                                                                              nxv, g.ny, g.nz, spacing, x_origin, gx);
    else
        voxelise_kernel<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(v, bytes, nxv, g.ny, g.nz, spacing, x_origin, gx);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (keep) *keep = bins; else bins.release();
    if (e != cudaSuccess) {
        if (keep) { keep->release(); }
        return lbm_internal_fail(LBM_ERR_CUDA, cudaGetErrorString(e));
    }
    return LBM_OK;
}

// Lattice::insertModel (Lattice.cpp:228-260): voxelise the mesh on the lattice points index*spacing.
// On a slab (lbm_set_slab_origin) the index is the global one and the neighbours' edge planes are voxelised too.
LBM_API int lbm_insert_model(lbm_handle h, lbm_mesh m, float spacing, int* accepted)
{
    int rc = lbm_internal_use(h);
    if (rc) return rc;
    if (!m) return lbm_internal_fail(LBM_ERR_INVALID, "null mesh");
    if (accepted) *accepted = 0;
    const Geom& g = h->g;
    if (!m->loaded || m->v.empty()) return LBM_OK;
    const bool slab = h->cfg.bc[0] == LBM_BC_HALO;
    if (slab && h->slab_gx <= 0) return lbm_internal_fail(LBM_ERR_STATE, "insert_model on a slab handle: call lbm_set_slab_origin first (or set the mask planes)");
    const int gx = slab ? h->slab_gx : g.nx;
    // the bounding-box guard: a mesh that leaves the box is silently ignored (Lattice.cpp:236-244)
    if (m->lo[0] < 0.0 || m->lo[1] < 0.0 || m->lo[2] < 0.0 || m->hi[0] > gx * spacing || m->hi[1] > g.ny * spacing ||
        m->hi[2] > g.nz * spacing)
        return LBM_OK;
    MeshView v;
    if ((rc = device_view(m, h->device, &v))) return rc;
    const size_t per = (size_t)g.ny * g.nz;
    const size_t n = (size_t)g.nx * per;
    if (!h->solid_bytes) GCK(cudaMalloc(&h->solid_bytes, n + 2 * per));
    if ((rc = slab ? voxelise_planes(h, m, v, h->solid_bytes, g.nx + 2, h->slab_x0 - 1, gx, spacing)
                   : voxelise_planes(h, m, v, h->solid_bytes, g.nx, 0, g.nx, spacing)))
        return rc;
    if (accepted) *accepted = 1;
    return lbm_internal_rebuild_mask(h, h->solid_bytes, slab ? 1 : 0);
}

// Lattice::preProcessModel (Lattice.cpp:323-426): build the reflection tables of the boundary-solid cells.
// On a slab the tables are built for its own planes and the neighbours' edge planes (a boundary-solid cell there
// reflects into this slab as well); the cells they look at reach one plane further, so the box that is voxelised for
// this spans two planes beyond each face (clipped to the lattice).
LBM_API int lbm_preprocess_model(lbm_handle h, lbm_mesh m, float spacing)
{
    int rc = lbm_internal_use(h);
    if (rc) return rc;
    if (!m) return lbm_internal_fail(LBM_ERR_INVALID, "null mesh");
    const Geom& g = h->g;
    const bool slab = h->cfg.bc[0] == LBM_BC_HALO;
    if (slab && h->slab_gx <= 0) return lbm_internal_fail(LBM_ERR_STATE, "preprocess_model on a slab handle: call lbm_set_slab_origin first");
    auto install = [&](const void* t, const int64_t* c, int cnt) {
        return slab ? lbm_set_reflection_table_slab(h, t, c, cnt) : lbm_set_reflection_table(h, t, c, cnt);
    };
    if (!h->solid || !h->any_solid) return install(nullptr, nullptr, 0);
    MeshView v;
    if ((rc = device_view(m, h->device, &v))) return rc;
    if (!m->loaded) v.nf = 0;
    // the box the tables are built on: the lattice itself, or the slab and two planes beyond each face
    const int xa = slab ? std::max(0, h->slab_x0 - 2) : 0;
    const int xb = slab ? std::min(h->slab_gx, h->slab_x0 + g.nx + 2) : g.nx;
    const int bx = xb - xa;
    const size_t per = (size_t)g.ny * g.nz;
    const size_t n = (size_t)bx * per;
    uint8_t* box = nullptr;          // solid flags of the box
    DeviceBins bins;                 // the triangles per lattice column of the box (first-hit searches look at two columns)
    if (slab) {
        GCK(cudaMalloc(&box, n));
        MeshView vv = v;
        if (!m->loaded) vv.nf = 0;
        if ((rc = voxelise_planes(h, m, vv, box, bx, xa, h->slab_gx, spacing, &bins))) { cudaFree(box); return rc; }
    } else {
        if (!h->solid_bytes) GCK(cudaMalloc(&h->solid_bytes, n + 2 * per));
        GCK(launch_mask_to_bytes(h->solid, h->solid_bytes, g, h->stream));
        box = h->solid_bytes;
        if ((rc = build_bins(h, m, bx, xa, spacing, &bins))) return rc;
    }
    auto drop_box = [&] { if (slab) cudaFree(box); bins.release(); };
    uint8_t* d_cand = nullptr;
    cudaError_t e = cudaMalloc(&d_cand, n);
    if (e != cudaSuccess) { drop_box(); return lbm_internal_fail(LBM_ERR_NOMEM, cudaGetErrorString(e)); }
    candidate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(box, d_cand, bx, g.ny, g.nz);
    std::vector<uint8_t> cand(n);
    e = cudaMemcpyAsync(cand.data(), d_cand, n, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_cand);
    if (e != cudaSuccess) { drop_box(); return lbm_internal_fail(LBM_ERR_CUDA, cudaGetErrorString(e)); }
    // candidates in host order = the order the reference discovers them in; a slab keeps those of its own planes and of
    // the neighbours' edge planes
    const int keep_lo = slab ? h->slab_x0 - 1 : 0, keep_hi = slab ? h->slab_x0 + g.nx : g.nx - 1;      // global planes, inclusive
    std::vector<long long> cells;
    for (size_t k = 0; k < n; ++k) {
        if (!cand[k]) continue;
        const int xg = (int)(k / per) + xa;
        if (xg >= keep_lo && xg <= keep_hi) cells.push_back((long long)k);
    }
    if (cells.empty()) { drop_box(); return install(nullptr, nullptr, 0); }
    long long* d_cells = nullptr;
    TableRecord* d_tables = nullptr;
    uint8_t* d_refl = nullptr;
    const int nc = (int)cells.size();
    e = cudaMalloc(&d_cells, nc * sizeof(long long));
    if (e == cudaSuccess) e = cudaMalloc(&d_tables, nc * sizeof(TableRecord));
    if (e == cudaSuccess) e = cudaMalloc(&d_refl, nc);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_cells, cells.data(), nc * sizeof(long long), cudaMemcpyHostToDevice, h->stream);
    std::vector<TableRecord> tables(nc);
    std::vector<uint8_t> refl(nc);
    if (e == cudaSuccess) {
        table_kernel<<<(nc + 63) / 64, 64, 0, h->stream>>>(v, box, d_cells, nc, d_tables, d_refl, bx, g.ny, g.nz, spacing, half_turn(), xa,
                                                           bins.ok ? bins.start : nullptr, bins.ok ? bins.tris : nullptr);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(tables.data(), d_tables, nc * sizeof(TableRecord), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(refl.data(), d_refl, nc, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_cells); cudaFree(d_tables); cudaFree(d_refl);
    drop_box();
    if (e != cudaSuccess) return lbm_internal_fail(LBM_ERR_CUDA, cudaGetErrorString(e));
    std::vector<TableRecord> keep;
    std::vector<int64_t> keep_cells;
    for (int c = 0; c < nc; ++c) {
        if (!refl[c]) continue;
        keep.push_back(tables[c]);
        // box index -> the handle's index space (a slab counts its planes from the low neighbour's edge plane)
        const long long shift = slab ? (long long)(xa - (h->slab_x0 - 1)) * (long long)per : 0;
        keep_cells.push_back(cells[c] + shift);
    }
    static_assert(sizeof(TableRecord) == 380, "ReflectionValues is 380 bytes");
    return install(keep.data(), keep_cells.data(), (int)keep.size());
}

// the tables currently set on the handle (by lbm_preprocess_model or lbm_set_reflection_table)
LBM_API int lbm_get_reflection_tables(lbm_handle h, void* tables, int64_t* cell_index, int capacity, int* n)
{
    if (!h || !n) return lbm_internal_fail(LBM_ERR_INVALID, "null argument");
    const int have = (int)h->host_table_cells.size();
    *n = have;
    if (!tables && !cell_index) return LBM_OK;
    if (capacity < have) return lbm_internal_fail(LBM_ERR_INVALID, "capacity too small");
    if (tables && have) std::memcpy(tables, h->host_tables.data(), (size_t)have * 380);
    if (cell_index && have) std::memcpy(cell_index, h->host_table_cells.data(), (size_t)have * sizeof(int64_t));
    return LBM_OK;
}

}  // extern "C"

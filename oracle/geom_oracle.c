/*
 * geom_oracle.c -- CPU restatement of the reference's one-off geometry preprocessing:
 * the side input (solid flags + reflection tables) that the per-timestep path consumes.
 * TEST INFRASTRUCTURE ONLY (see lbm_oracle.h).  Citations relative to /root/reference.
 *
 * Third-party arithmetic: the reference answers "is this point inside the mesh" and
 * "which face does this ray hit first" with CGAL (Surface_mesh, AABB_tree,
 * Side_of_triangle_mesh, PMP::compute_face_normal over Simple_cartesian<float>;
 * libcgal-dev from Ubuntu 22.04 apt, version unpinned -- Dockerfile:5-9, CMakeLists.txt:8).
 * CGAL is not vendored and not installed here, so those two queries restate CGAL's
 * published algorithm (ray casting with re-shooting on degenerate hits; a point lying on
 * a facet is ON_BOUNDARY) and are pinned only on the reference's own fixtures and call
 * sites: test/ModelTest.cpp:5-74, test/CoordinatesTest.cpp:137-229,
 * test/CalculationsTest.cpp:107-197.  Everything downstream of those two queries
 * (ranking, slot packing, weights, bounding-box guard) is a line-for-line restatement.
 */
#include "lbm_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <ctype.h>

struct lbo_mesh {
    int nv, nf;
    float *v;        /* nv * 3 */
    int *face_start; /* nf + 1 */
    int *face_idx;   /* concatenated vertex indices */
    int loaded;      /* 0: file was missing (reference returns early, Model.cpp:24-27) */
    int valid;       /* Model::meshValid */
};

static void *grow(void *p, size_t *cap, size_t need, size_t elem)
{
    if (need <= *cap) return p;
    size_t c = *cap ? *cap * 2 : 64;
    while (c < need) c *= 2;
    *cap = c;
    return realloc(p, c * elem);
}

int lbo_mesh_num_vertices(const lbo_mesh *m) { return m->nv; }
int lbo_mesh_num_faces(const lbo_mesh *m) { return m->nf; }
int lbo_mesh_check_error(const lbo_mesh *m) { return !m->valid; }

static int mesh_is_triangle(const lbo_mesh *m)
{
    for (int f = 0; f < m->nf; ++f)
        if (m->face_start[f + 1] - m->face_start[f] != 3) return 0;
    return 1;
}

static int cmp_ll(const void *pa, const void *pb)
{
    long long a = *(const long long *)pa, b = *(const long long *)pb;
    return (a > b) - (a < b);
}

/* CGAL::is_closed: no border halfedge, i.e. every undirected edge is used by two faces */
int lbo_mesh_is_closed(const lbo_mesh *m)
{
    int ne = m->face_start[m->nf];
    if (ne == 0) return 1;
    long long *keys = (long long *)malloc(sizeof(long long) * (size_t)ne);
    int k = 0;
    for (int f = 0; f < m->nf; ++f) {
        int b = m->face_start[f], e = m->face_start[f + 1];
        for (int j = b; j < e; ++j) {
            int a0 = m->face_idx[j], a1 = m->face_idx[(j + 1 < e) ? j + 1 : b];
            int lo = a0 < a1 ? a0 : a1, hi = a0 < a1 ? a1 : a0;
            keys[k++] = (long long)lo * (long long)(m->nv + 1) + hi;
        }
    }
    qsort(keys, (size_t)ne, sizeof(long long), cmp_ll);
    int closed = 1;
    for (int i = 0; i < ne && closed; ) {
        int j = i;
        while (j < ne && keys[j] == keys[i]) ++j;
        if (j - i != 2) closed = 0;
        i = j;
    }
    free(keys);
    return closed;
}

lbo_mesh *lbo_mesh_load(const char *path)
{
    lbo_mesh *m = (lbo_mesh *)calloc(1, sizeof(*m));
    if (!m) return NULL;
    m->valid = 1;
    m->face_start = (int *)calloc(1, sizeof(int));
    FILE *fp = fopen(path, "r");
    if (!fp) return m;                      /* Model.cpp:24-27: silently keep an empty model */
    m->loaded = 1;
    size_t vcap = 0, fcap = 1, icap = 0;
    int nidx = 0;
    char line[4096];
    while (fgets(line, sizeof line, fp)) {
        char *hash = strchr(line, '#');
        if (hash) *hash = 0;
        char *p = line;
        while (isspace((unsigned char)*p)) ++p;
        if (p[0] == 'v' && isspace((unsigned char)p[1])) {
            double a, b, c;
            if (sscanf(p + 1, "%lf %lf %lf", &a, &b, &c) == 3) {
                m->v = (float *)grow(m->v, &vcap, (size_t)(m->nv + 1) * 3, sizeof(float));
                m->v[m->nv * 3 + 0] = (float)a; m->v[m->nv * 3 + 1] = (float)b; m->v[m->nv * 3 + 2] = (float)c;
                ++m->nv;
            }
        } else if (p[0] == 'f' && isspace((unsigned char)p[1])) {
            p += 1;
            int count = 0;
            for (;;) {
                while (isspace((unsigned char)*p)) ++p;
                if (!*p) break;
                char *end;
                long id = strtol(p, &end, 10);
                if (end == p) break;
                while (*end && !isspace((unsigned char)*end)) ++end;   /* skip /vt/vn */
                p = end;
                int vi = (id > 0) ? (int)id - 1 : m->nv + (int)id;
                m->face_idx = (int *)grow(m->face_idx, &icap, (size_t)nidx + 1, sizeof(int));
                m->face_idx[nidx++] = vi;
                ++count;
            }
            if (count >= 3) {
                m->face_start = (int *)grow(m->face_start, &fcap, (size_t)m->nf + 2, sizeof(int));
                ++m->nf;
                m->face_start[m->nf] = nidx;
            } else {
                nidx -= count;
            }
        }
    }
    fclose(fp);
    m->valid = mesh_is_triangle(m);          /* Model.cpp:31 */
    if (!lbo_mesh_is_closed(m)) m->valid = 0; /* Model.cpp:40-43 */
    return m;
}

void lbo_mesh_free(lbo_mesh *m)
{
    if (!m) return;
    free(m->v); free(m->face_start); free(m->face_idx); free(m);
}

void lbo_mesh_bbox(const lbo_mesh *m, float lo[3], float hi[3])
{
    for (int a = 0; a < 3; ++a) { lo[a] = 0.0f; hi[a] = 0.0f; }
    for (int i = 0; i < m->nv; ++i)
        for (int a = 0; a < 3; ++a) {
            float c = m->v[i * 3 + a];
            if (i == 0 || c < lo[a]) lo[a] = c;
            if (i == 0 || c > hi[a]) hi[a] = c;
        }
}

/* ---- ray / triangle predicates (double arithmetic on float inputs) --------------- */

static void tri(const lbo_mesh *m, int f, double a[3], double b[3], double c[3])
{
    const int *id = &m->face_idx[m->face_start[f]];
    for (int k = 0; k < 3; ++k) { a[k] = m->v[id[0] * 3 + k]; b[k] = m->v[id[1] * 3 + k]; c[k] = m->v[id[2] * 3 + k]; }
}

static void cross3(const double a[3], const double b[3], double o[3])
{
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
static double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

/* 1 if p lies on the closed triangle f */
static int point_on_triangle(const lbo_mesh *m, int f, const double p[3])
{
    double a[3], b[3], c[3], e1[3], e2[3], n[3], d[3];
    tri(m, f, a, b, c);
    for (int k = 0; k < 3; ++k) { e1[k] = b[k] - a[k]; e2[k] = c[k] - a[k]; d[k] = p[k] - a[k]; }
    cross3(e1, e2, n);
    double nn = dot3(n, n);
    if (nn == 0.0) return 0;
    double dist = dot3(n, d);
    double scale = sqrt(nn) * (sqrt(dot3(d, d)) + 1e-30);
    if (fabs(dist) > 1e-12 * scale) return 0;          /* not coplanar */
    /* barycentric coordinates in the plane */
    double d11 = dot3(e1, e1), d12 = dot3(e1, e2), d22 = dot3(e2, e2), p1 = dot3(d, e1), p2 = dot3(d, e2);
    double det = d11 * d22 - d12 * d12;
    double u = (d22 * p1 - d12 * p2) / det, v = (d11 * p2 - d12 * p1) / det;
    const double eps = 1e-9;
    return u >= -eps && v >= -eps && u + v <= 1.0 + eps;
}

/* ray o + t*dir vs triangle f.  returns 0 miss, 1 clean interior hit, 2 degenerate
 * (parallel-in-plane, or through an edge/vertex) -- the caller re-shoots on 2. */
static int ray_triangle(const lbo_mesh *m, int f, const double o[3], const double dir[3], double *t_out)
{
    double a[3], b[3], c[3], e1[3], e2[3], pv[3], tv[3], qv[3];
    tri(m, f, a, b, c);
    for (int k = 0; k < 3; ++k) { e1[k] = b[k] - a[k]; e2[k] = c[k] - a[k]; tv[k] = o[k] - a[k]; }
    cross3(dir, e2, pv);
    double det = dot3(e1, pv);
    double scale = sqrt(dot3(e1, e1) * dot3(e2, e2) * dot3(dir, dir)) + 1e-300;
    const double eps = 1e-9;
    if (fabs(det) <= 1e-12 * scale) {
        /* ray parallel to the facet plane: degenerate only if it lies in the plane */
        double n[3]; cross3(e1, e2, n);
        double off = dot3(n, tv);
        if (fabs(off) <= 1e-12 * sqrt(dot3(n, n)) * (sqrt(dot3(tv, tv)) + 1e-30)) return 2;
        return 0;
    }
    double inv = 1.0 / det;
    double u = dot3(tv, pv) * inv;
    cross3(tv, e1, qv);
    double v = dot3(dir, qv) * inv;
    double t = dot3(e2, qv) * inv;
    if (u < -eps || v < -eps || u + v > 1.0 + eps) return 0;
    if (t < -eps) return 0;
    *t_out = t;
    if (u < eps || v < eps || u + v > 1.0 - eps || t < eps) return 2;
    return 1;
}

/* Model::isPointInsideModel, Model.cpp:51-63: ON_BOUNDED_SIDE or ON_BOUNDARY => true.
 * Restates CGAL::Side_of_triangle_mesh (facet test first, then a vertical ray, then
 * re-shot rays while a ray grazes an edge or vertex). */
int lbo_mesh_point_inside(const lbo_mesh *m, float px, float py, float pz)
{
    if (!m->valid || !m->loaded || m->nf == 0) return 0;
    const double p[3] = {px, py, pz};
    float lo[3], hi[3];
    lbo_mesh_bbox(m, lo, hi);
    if (px < lo[0] || px > hi[0] || py < lo[1] || py > hi[1] || pz < lo[2] || pz > hi[2]) return 0;
    for (int f = 0; f < m->nf; ++f)
        if (point_on_triangle(m, f, p)) return 1;      /* ON_BOUNDARY */
    unsigned long long s = 0x9E3779B97F4A7C15ull;
    for (int attempt = 0; attempt < 64; ++attempt) {
        double dir[3];
        if (attempt == 0) { dir[0] = 0.0; dir[1] = 0.0; dir[2] = 1.0; }
        else for (int k = 0; k < 3; ++k) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            dir[k] = ((double)(s >> 11) / 9007199254740992.0) * 2.0 - 1.0;
        }
        int crossings = 0, bad = 0;
        for (int f = 0; f < m->nf && !bad; ++f) {
            double t;
            int r = ray_triangle(m, f, p, dir, &t);
            if (r == 2) bad = 1; else if (r == 1) ++crossings;
        }
        if (!bad) return crossings & 1;
    }
    return 0;
}

/* calculateVectorMagnitude + quaternionMult + rotateVectorAroundAxis, Geometry.cpp:7-42.
 * float arithmetic throughout; the half angle is formed in double then narrowed. */
static void quat_mul(const float a[4], const float b[4], float r[4])
{
    float r0 = (a[0] * b[0]) - (a[1] * b[1]) - (a[2] * b[2]) - (a[3] * b[3]);
    float r1 = (a[0] * b[1]) + (a[1] * b[0]) + (a[2] * b[3]) - (a[3] * b[2]);
    float r2 = (a[0] * b[2]) - (a[1] * b[3]) + (a[2] * b[0]) + (a[3] * b[1]);
    float r3 = (a[0] * b[3]) + (a[1] * b[2]) - (a[2] * b[1]) + (a[3] * b[0]);
    r[0] = r0; r[1] = r1; r[2] = r2; r[3] = r3;
}

void lbo_rotate_vector_around_axis(const float axis_in[3], const float vec[3], float angle, float out[3])
{
    float axis[3] = {axis_in[0], axis_in[1], axis_in[2]};
    float magnitude = sqrtf((axis[0] * axis[0]) + (axis[1] * axis[1]) + (axis[2] * axis[2]));
    axis[0] /= magnitude; axis[1] /= magnitude; axis[2] /= magnitude;
    angle = (float)(angle / 2.0);
    float q[4]  = {cosf(angle),  sinf(angle)  * axis[0], sinf(angle)  * axis[1], sinf(angle)  * axis[2]};
    float qi[4] = {cosf(-angle), sinf(-angle) * axis[0], sinf(-angle) * axis[1], sinf(-angle) * axis[2]};
    float pt[4] = {0.0f, vec[0], vec[1], vec[2]};
    float tmp[4], res[4];
    quat_mul(q, pt, tmp);
    quat_mul(tmp, qi, res);
    out[0] = res[1]; out[1] = res[2]; out[2] = res[3];
}

/* Model::reflectionVector, Model.cpp:75-91 */
void lbo_mesh_reflection_vector(const lbo_mesh *m, const float p1[3], const float p2[3], float out[3])
{
    out[0] = out[1] = out[2] = 0.0f;
    if (!m->loaded || m->nf == 0) return;
    const double o[3] = {p1[0], p1[1], p1[2]};
    const double dir[3] = {(double)p2[0] - p1[0], (double)p2[1] - p1[1], (double)p2[2] - p1[2]};
    /* AABB_tree::first_intersection(Ray_3(point_1, point_2)): nearest facet along the ray */
    int best = -1; double best_t = 0.0;
    for (int f = 0; f < m->nf; ++f) {
        if (m->face_start[f + 1] - m->face_start[f] != 3) continue;
        double t;
        int r = ray_triangle(m, f, o, dir, &t);
        if (r == 0) continue;
        if (r == 2) {
            /* in-plane rays do not give a usable hit; edge/vertex grazes do */
            double a[3], b[3], c[3], e1[3], e2[3], n[3];
            tri(m, f, a, b, c);
            for (int k = 0; k < 3; ++k) { e1[k] = b[k] - a[k]; e2[k] = c[k] - a[k]; }
            cross3(e1, e2, n);
            if (fabs(dot3(n, dir)) <= 1e-12 * sqrt(dot3(n, n) * dot3(dir, dir))) continue;
        }
        if (best < 0 || t < best_t - 1e-12) { best = f; best_t = t; }
    }
    /* Model.cpp:79-81: vector = normalise(point_1 - point_2) in float */
    float vx = p1[0] - p2[0], vy = p1[1] - p2[1], vz = p1[2] - p2[2];
    float norm = sqrtf((vx * vx) + (vy * vy) + (vz * vz));
    vx /= norm; vy /= norm; vz /= norm;
    if (best < 0) return;                                   /* Model.cpp:90 */
    /* PMP::compute_face_normal over Simple_cartesian<float>: unit cross product */
    const int *id = &m->face_idx[m->face_start[best]];
    const float *A = &m->v[id[0] * 3], *B = &m->v[id[1] * 3], *C = &m->v[id[2] * 3];
    float e1[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
    float e2[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
    float n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
    float nl = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    if (nl != 0.0f) { n[0] /= nl; n[1] /= nl; n[2] /= nl; }
    float v[3] = {vx, vy, vz};
    lbo_rotate_vector_around_axis(n, v, (float)M_PI, out);   /* Model.cpp:86 */
}

/* Lattice::distributeVector, Lattice.cpp:262-321 */
void lbo_distribute_vector(const float v[3], float scores[27], int dirs[27])
{
    float unit[27][3];
    for (int i = 0; i < 27; ++i) {
        int nz = abs(LBO_DIRECTIONS[i][0]) + abs(LBO_DIRECTIONS[i][1]) + abs(LBO_DIRECTIONS[i][2]);
        double len = (nz == 2) ? 1.4142135623730951 : (nz == 3) ? 1.7320508075688772 : 1.0;
        for (int a = 0; a < 3; ++a)
            unit[i][a] = (LBO_DIRECTIONS[i][a] == 0) ? 0.0f : (float)((double)LBO_DIRECTIONS[i][a] / len);
    }
    float sc[27];
    for (int i = 0; i < 27; ++i) {
        float dx = v[0] - unit[i][0], dy = v[1] - unit[i][1], dz = v[2] - unit[i][2];
        float d = sqrtf((dx * dx) + (dy * dy) + (dz * dz));
        sc[i] = (float)(2.0 - d);                              /* Lattice.cpp:296 */
    }
    /* max-heap of (score, index) pairs popped 27 times = descending lexicographic order */
    int used[27] = {0};
    for (int r = 0; r < 27; ++r) {
        int best = -1;
        for (int i = 0; i < 27; ++i) {
            if (used[i]) continue;
            if (best < 0 || sc[i] > sc[best] || (sc[i] == sc[best] && i > best)) best = i;
        }
        used[best] = 1; scores[r] = sc[best]; dirs[r] = best;
    }
}

int lbo_insert_model(const lbo_mesh *m, lbo_lattice_point *cells, int X, int Y, int Z, float spacing)
{
    if (!m->loaded || m->nv == 0) return 0;
    float lo[3], hi[3];
    lbo_mesh_bbox(m, lo, hi);
    /* Lattice.cpp:236-244 */
    if (lo[0] < 0.0 || lo[1] < 0.0 || lo[2] < 0.0 ||
        hi[0] > X * spacing || hi[1] > Y * spacing || hi[2] > Z * spacing)
        return 0;
    /* (every cell is an independent query: the loop is spread over the cores, the answers do not depend on that) */
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int x = 0; x < X; ++x)
        for (int y = 0; y < Y; ++y)
            for (int z = 0; z < Z; ++z)
                cells[lbo_index(x, y, z, Y, Z)].is_internal =
                    (uint8_t)lbo_mesh_point_inside(m, x * spacing, y * spacing, z * spacing);   /* :254 */
    return 1;
}

int lbo_preprocess_model(const lbo_mesh *m, lbo_lattice_point *cells, int X, int Y, int Z, float spacing,
                         lbo_reflection_values *tables, int max_tables, int32_t *table_index)
{
    size_t n = (size_t)X * Y * Z;
    for (size_t k = 0; k < n; ++k) table_index[k] = -1;
    int ntab = 0;
    for (int x = 0; x < X; ++x)
        for (int y = 0; y < Y; ++y)
            for (int z = 0; z < Z; ++z) {
                size_t k = lbo_index(x, y, z, Y, Z);
                lbo_lattice_point *cur = &cells[k];
                if (!cur->is_internal) continue;                                   /* :342 */
                for (int d = 0; d < 27; ++d) {
                    int nx = x + LBO_DIRECTIONS[d][0], ny = y + LBO_DIRECTIONS[d][1], nz = z + LBO_DIRECTIONS[d][2];
                    if (nx < 0 || ny < 0 || nz < 0 || nx >= X || ny >= Y || nz >= Z) continue;   /* :346-354 */
                    if (cells[lbo_index(nx, ny, nz, Y, Z)].is_internal) continue;             /* :356 */
                    float p1[3] = {spacing * nx, spacing * ny, spacing * nz};
                    float p2[3] = {spacing * x, spacing * y, spacing * z};
                    float r[3];
                    lbo_mesh_reflection_vector(m, p1, p2, r);                       /* :358-359 */
                    float scores[27]; int dirs[27];
                    lbo_distribute_vector(r, scores, dirs);                         /* :361 */
                    int counter = 0;
                    float norm = 0.0;
                    int base = LBO_INVERSE[d] * 3;
                    for (int i = 0; i < 27 && counter < 3; ++i) {
                        int c = dirs[i];
                        int tx = x + LBO_DIRECTIONS[c][0], ty = y + LBO_DIRECTIONS[c][1], tz = z + LBO_DIRECTIONS[c][2];
                        /* the reference indexes without a bounds check here (:369);
                         * an out-of-domain neighbour is treated as unavailable */
                        if (tx < 0 || ty < 0 || tz < 0 || tx >= X || ty >= Y || tz >= Z) continue;
                        if (cells[lbo_index(tx, ty, tz, Y, Z)].is_internal) continue;
                        if (table_index[k] < 0) {                                   /* :371-382 */
                            if (ntab >= max_tables) return -1;
                            memset(&tables[ntab], 0, sizeof(tables[ntab]));
                            table_index[k] = ntab++;
                        }
                        lbo_reflection_values *t = &tables[table_index[k]];
                        lbo_set_reflection_direction(t, base + counter, (unsigned)c);   /* :384 */
                        t->reflection_weight[base + counter] = scores[i];           /* :385 */
                        norm += scores[i];
                        cur->is_reflected = 1;
                        ++counter;
                    }
                    if (norm != 0.0) {                                              /* :400-409 */
                        lbo_reflection_values *t = &tables[table_index[k]];
                        for (int i = 0; i < 3; ++i)
                            if (lbo_get_reflection_direction(t, base + i) != 0)
                                t->reflection_weight[base + i] /= norm;
                    }
                }
            }
    return ntab;
}

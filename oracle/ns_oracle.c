/* ns_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE). See ns_oracle.h.
 *
 * PARITY UNPINNED (no golden vectors exist in the reference; ugcore is absent).
 * Every routine cites the reference lines it restates (paths relative to /root/reference).
 * Written as straightforward serial C so that it reads like the reference's element
 * routines; the product's CUDA kernels are an independent implementation.
 */
#include "ns_oracle.h"
#include <math.h>
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXSH 8
#define MAXIP 12
#define MAXL  32

static __thread char g_err[512];
static char g_err_shared[512];
const char *ora_last_error(void) { return g_err_shared; }
static int fail(const char *msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
#pragma omp critical(ora_err)
    { snprintf(g_err_shared, sizeof g_err_shared, "%s", msg); }
    return -1;
}

/* ------------------------------------------------------------------------------------------
 * Reference elements (ugcore lib_disc/reference_element -- our spec)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int dim, nsh, nedge, nside;
    double corner[MAXSH][3];
    int edge[MAXIP][2];
    int side_n[6];
    int side[6][4];
    /* FV1 sub-control-volume faces: one per edge. 3-D: faces A/B adjacent to the edge,
       ordered so that 0.5*(c2-c0)x(c3-c1) points from edge corner 0 to edge corner 1. */
    int scvf_faceA[MAXIP], scvf_faceB[MAXIP];
    double lip[MAXIP][3];
    double shape_ip[MAXIP][MAXSH];
    double lgrad_ip[MAXIP][MAXSH][3];
    int ready;
} RefElem;

static RefElem g_ref[5];

static const double TRI_CO[3][3]  = {{0,0,0},{1,0,0},{0,1,0}};
static const double QUAD_CO[4][3] = {{0,0,0},{1,0,0},{1,1,0},{0,1,0}};
static const double TET_CO[4][3]  = {{0,0,0},{1,0,0},{0,1,0},{0,0,1}};
static const double HEX_CO[8][3]  = {{0,0,0},{1,0,0},{1,1,0},{0,1,0},{0,0,1},{1,0,1},{1,1,1},{0,1,1}};
static const int TRI_ED[3][2]  = {{0,1},{1,2},{2,0}};
static const int QUAD_ED[4][2] = {{0,1},{1,2},{2,3},{3,0}};
static const int TET_ED[6][2]  = {{0,1},{1,2},{2,0},{0,3},{1,3},{2,3}};
static const int HEX_ED[12][2] = {{0,1},{1,2},{2,3},{3,0},{0,4},{1,5},{2,6},{3,7},{4,5},{5,6},{6,7},{7,4}};
static const int TET_FA[4][4]  = {{0,2,1,-1},{1,2,3,-1},{0,3,2,-1},{0,1,3,-1}};
static const int HEX_FA[6][4]  = {{0,3,2,1},{0,1,5,4},{1,2,6,5},{2,3,7,6},{3,0,4,7},{4,5,6,7}};
/* prism (ugcore ReferencePrism: bottom triangle 0,1,2, top triangle 3,4,5; sides: bottom, three quadrilaterals, top) */
static const double PRISM_CO[6][3] = {{0,0,0},{1,0,0},{0,1,0},{0,0,1},{1,0,1},{0,1,1}};
static const int PRISM_ED[9][2] = {{0,1},{1,2},{2,0},{0,3},{1,4},{2,5},{3,4},{4,5},{5,3}};
static const int PRISM_FA[5][4] = {{0,2,1,-1},{0,1,4,3},{1,2,5,4},{2,0,3,5},{3,4,5,-1}};
static const int PRISM_FN[5]    = {3,4,4,4,3};

int ora_elem_nsh(int e)  { static const int v[5] = {3,4,4,8,6};  return (e>=0&&e<5)? v[e] : -1; }
int ora_elem_nip(int e)  { static const int v[5] = {3,4,6,12,9}; return (e>=0&&e<5)? v[e] : -1; }
int ora_elem_dim(int e)  { static const int v[5] = {2,2,3,3,3};  return (e>=0&&e<5)? v[e] : -1; }
int ora_elem_nside(int e){ static const int v[5] = {3,4,4,6,5};  return (e>=0&&e<5)? v[e] : -1; }

/* P1/Q1 Lagrange shapes and local gradients (ugcore LagrangeP1<RefElem>) */
static void lagrange_shapes(int elem, const double *xi, double *N, double (*dN)[3])
{
    double x = xi[0], y = xi[1], z = xi[2];
    int i;
    switch (elem) {
    case ORA_TRI:
        N[0] = 1-x-y; N[1] = x; N[2] = y;
        if (dN) { dN[0][0]=-1; dN[0][1]=-1; dN[1][0]=1; dN[1][1]=0; dN[2][0]=0; dN[2][1]=1;
                  for(i=0;i<3;i++) dN[i][2]=0; }
        break;
    case ORA_QUAD:
        N[0]=(1-x)*(1-y); N[1]=x*(1-y); N[2]=x*y; N[3]=(1-x)*y;
        if (dN) { dN[0][0]=-(1-y); dN[0][1]=-(1-x); dN[1][0]=(1-y); dN[1][1]=-x;
                  dN[2][0]=y; dN[2][1]=x; dN[3][0]=-y; dN[3][1]=(1-x);
                  for(i=0;i<4;i++) dN[i][2]=0; }
        break;
    case ORA_TET:
        N[0]=1-x-y-z; N[1]=x; N[2]=y; N[3]=z;
        if (dN) { dN[0][0]=-1;dN[0][1]=-1;dN[0][2]=-1; dN[1][0]=1;dN[1][1]=0;dN[1][2]=0;
                  dN[2][0]=0;dN[2][1]=1;dN[2][2]=0; dN[3][0]=0;dN[3][1]=0;dN[3][2]=1; }
        break;
    case ORA_HEX:
        for (i = 0; i < 8; i++) {
            double sx = HEX_CO[i][0] > 0.5 ? 1.0 : -1.0, fx = HEX_CO[i][0] > 0.5 ? x : 1-x;
            double sy = HEX_CO[i][1] > 0.5 ? 1.0 : -1.0, fy = HEX_CO[i][1] > 0.5 ? y : 1-y;
            double sz = HEX_CO[i][2] > 0.5 ? 1.0 : -1.0, fz = HEX_CO[i][2] > 0.5 ? z : 1-z;
            N[i] = fx*fy*fz;
            if (dN) { dN[i][0]=sx*fy*fz; dN[i][1]=fx*sy*fz; dN[i][2]=fx*fy*sz; }
        }
        break;
    case ORA_PRISM: {
        /* P1 on the triangle times P1 along the axis (ugcore LagrangeP1<ReferencePrism>) */
        const double l[3] = {1-x-y, x, y}, dl[3][2] = {{-1,-1},{1,0},{0,1}};
        for (i = 0; i < 6; i++) {
            const int t = i % 3; const double fz = i < 3 ? 1-z : z, sz = i < 3 ? -1.0 : 1.0;
            N[i] = l[t]*fz;
            if (dN) { dN[i][0]=dl[t][0]*fz; dN[i][1]=dl[t][1]*fz; dN[i][2]=l[t]*sz; }
        }
        break; }
    }
}

static void vcross(double *o, const double *a, const double *b)
{ o[0]=a[1]*b[2]-a[2]*b[1]; o[1]=a[2]*b[0]-a[0]*b[2]; o[2]=a[0]*b[1]-a[1]*b[0]; }
static double vdot(const double *a, const double *b, int dim)
{ double s = 0; for (int d = 0; d < dim; d++) s += a[d]*b[d]; return s; }
static double vdist(const double *a, const double *b, int dim)
{ double s = 0; for (int d = 0; d < dim; d++) s += (a[d]-b[d])*(a[d]-b[d]); return sqrt(s); }
static double vdistsq(const double *a, const double *b, int dim)
{ double s = 0; for (int d = 0; d < dim; d++) s += (a[d]-b[d])*(a[d]-b[d]); return s; }

/* average of a subset of points */
static void avg_pts(double *o, const double (*x)[3], const int *ids, int n, int dim)
{
    for (int d = 0; d < 3; d++) o[d] = 0;
    for (int i = 0; i < n; i++) for (int d = 0; d < dim; d++) o[d] += x[ids[i]][d];
    for (int d = 0; d < dim; d++) o[d] /= n;
}

/* SCVF corner positions from element corner positions x (local or global), App. B-2:
   2-D [edge midpoint, barycentre]; 3-D [edge midpoint, centre face A, barycentre, centre face B] */
static int scvf_corners(const RefElem *r, int ip, const double (*x)[3], double (*c)[3])
{
    int all[MAXSH]; for (int i = 0; i < r->nsh; i++) all[i] = i;
    avg_pts(c[0], x, r->edge[ip], 2, r->dim);
    if (r->dim == 2) { avg_pts(c[1], x, all, r->nsh, 2); return 2; }
    avg_pts(c[1], x, r->side[r->scvf_faceA[ip]], r->side_n[r->scvf_faceA[ip]], 3);
    avg_pts(c[2], x, all, r->nsh, 3);
    avg_pts(c[3], x, r->side[r->scvf_faceB[ip]], r->side_n[r->scvf_faceB[ip]], 3);
    return 4;
}

/* area-scaled SCVF normal (ugcore NormalOnSCVF): 2-D (dy,-dx) of c1-c0; 3-D 0.5*(c2-c0)x(c3-c1) */
static void scvf_normal(int dim, const double (*c)[3], double *n)
{
    if (dim == 2) { n[0] = c[1][1]-c[0][1]; n[1] = -(c[1][0]-c[0][0]); n[2] = 0; return; }
    double a[3], b[3];
    for (int d = 0; d < 3; d++) { a[d] = c[2][d]-c[0][d]; b[d] = c[3][d]-c[1][d]; }
    vcross(n, a, b);
    for (int d = 0; d < 3; d++) n[d] *= 0.5;
}

static const RefElem *get_ref(int elem)
{
    if (elem < 0 || elem > 4) return NULL;
    RefElem *r = &g_ref[elem];
    if (r->ready) return r;
#pragma omp critical(ora_ref_init)
    if (!r->ready) {
        RefElem t; memset(&t, 0, sizeof t);
        t.dim = ora_elem_dim(elem); t.nsh = ora_elem_nsh(elem); t.nedge = ora_elem_nip(elem);
        t.nside = ora_elem_nside(elem);
        for (int i = 0; i < t.nsh; i++) for (int d = 0; d < 3; d++)
            t.corner[i][d] = elem==ORA_TRI ? TRI_CO[i][d] : elem==ORA_QUAD ? QUAD_CO[i][d]
                           : elem==ORA_TET ? TET_CO[i][d] : elem==ORA_HEX ? HEX_CO[i][d] : PRISM_CO[i][d];
        for (int i = 0; i < t.nedge; i++) for (int k = 0; k < 2; k++)
            t.edge[i][k] = elem==ORA_TRI ? TRI_ED[i][k] : elem==ORA_QUAD ? QUAD_ED[i][k]
                         : elem==ORA_TET ? TET_ED[i][k] : elem==ORA_HEX ? HEX_ED[i][k] : PRISM_ED[i][k];
        for (int s = 0; s < t.nside; s++) {
            if (t.dim == 2) { t.side_n[s] = 2; t.side[s][0] = t.edge[s][0]; t.side[s][1] = t.edge[s][1]; }
            else if (elem == ORA_TET) { t.side_n[s] = 3; for (int k=0;k<3;k++) t.side[s][k] = TET_FA[s][k]; }
            else if (elem == ORA_HEX) { t.side_n[s] = 4; for (int k=0;k<4;k++) t.side[s][k] = HEX_FA[s][k]; }
            else { t.side_n[s] = PRISM_FN[s]; for (int k=0;k<4;k++) t.side[s][k] = PRISM_FA[s][k]; }
        }
        for (int ip = 0; ip < t.nedge; ip++) {
            if (t.dim == 3) {
                int f[2], nf = 0;
                for (int s = 0; s < t.nside && nf < 2; s++) {
                    int h0 = 0, h1 = 0;
                    for (int k = 0; k < t.side_n[s]; k++) {
                        if (t.side[s][k] == t.edge[ip][0]) h0 = 1;
                        if (t.side[s][k] == t.edge[ip][1]) h1 = 1;
                    }
                    if (h0 && h1) f[nf++] = s;
                }
                t.scvf_faceA[ip] = f[0]; t.scvf_faceB[ip] = f[1];
                double c[4][3], n[3], e[3];
                scvf_corners(&t, ip, t.corner, c); scvf_normal(3, c, n);
                for (int d = 0; d < 3; d++) e[d] = t.corner[t.edge[ip][1]][d] - t.corner[t.edge[ip][0]][d];
                if (vdot(n, e, 3) < 0) { t.scvf_faceA[ip] = f[1]; t.scvf_faceB[ip] = f[0]; }
            }
            double c[4][3]; int nc = scvf_corners(&t, ip, t.corner, c);
            for (int d = 0; d < 3; d++) { t.lip[ip][d] = 0; for (int k = 0; k < nc; k++) t.lip[ip][d] += c[k][d]; t.lip[ip][d] /= nc; }
            lagrange_shapes(elem, t.lip[ip], t.shape_ip[ip], t.lgrad_ip[ip]);
        }
        t.ready = 1;
        *r = t;
    }
    return r;
}

/* inverse of a dim x dim matrix (row-major 3x3 storage). returns det */
static double mat_inverse(int dim, const double a[3][3], double inv[3][3])
{
    if (dim == 2) {
        double det = a[0][0]*a[1][1]-a[0][1]*a[1][0];
        inv[0][0] =  a[1][1]/det; inv[0][1] = -a[0][1]/det;
        inv[1][0] = -a[1][0]/det; inv[1][1] =  a[0][0]/det;
        return det;
    }
    double c00 = a[1][1]*a[2][2]-a[1][2]*a[2][1];
    double c01 = a[1][2]*a[2][0]-a[1][0]*a[2][2];
    double c02 = a[1][0]*a[2][1]-a[1][1]*a[2][0];
    double det = a[0][0]*c00 + a[0][1]*c01 + a[0][2]*c02;
    inv[0][0] = c00/det; inv[1][0] = c01/det; inv[2][0] = c02/det;
    inv[0][1] = (a[0][2]*a[2][1]-a[0][1]*a[2][2])/det;
    inv[1][1] = (a[0][0]*a[2][2]-a[0][2]*a[2][0])/det;
    inv[2][1] = (a[0][1]*a[2][0]-a[0][0]*a[2][1])/det;
    inv[0][2] = (a[0][1]*a[1][2]-a[0][2]*a[1][1])/det;
    inv[1][2] = (a[0][2]*a[1][0]-a[0][0]*a[1][2])/det;
    inv[2][2] = (a[0][0]*a[1][1]-a[0][1]*a[1][0])/det;
    return det;
}

/* exact volume of a trilinear hexahedron, corners in reference order (long-diagonal formula) */
static double hex_volume(const double (*p)[3])
{
    double a[3], b[3], c[3], t[3], v = 0;
    /* V = 1/12 * ( [(p6-p1)+(p7-p0), p6-p3, p2-p0] + [p7-p0, (p6-p3)+(p5-p0), p6-p4]
                  + [p6-p1, p5-p0, (p6-p4)+(p2-p0)] ) */
    for (int d=0; d<3; d++) { a[d]=(p[6][d]-p[1][d])+(p[7][d]-p[0][d]); b[d]=p[6][d]-p[3][d]; c[d]=p[2][d]-p[0][d]; }
    vcross(t, b, c); v += vdot(a, t, 3);
    for (int d=0; d<3; d++) { a[d]=p[7][d]-p[0][d]; b[d]=(p[6][d]-p[3][d])+(p[5][d]-p[0][d]); c[d]=p[6][d]-p[4][d]; }
    vcross(t, b, c); v += vdot(a, t, 3);
    for (int d=0; d<3; d++) { a[d]=p[6][d]-p[1][d]; b[d]=p[5][d]-p[0][d]; c[d]=(p[6][d]-p[4][d])+(p[2][d]-p[0][d]); }
    vcross(t, b, c); v += vdot(a, t, 3);
    return v / 12.0;
}

/* ------------------------------------------------------------------------------------------
 * FV1Geometry::update restatement (ugcore fv1_geom.cpp -- SURVEY App. B-2, our spec):
 * one SCV per corner, one SCVF per edge; ip = mean of SCVF corners; shapes/gradients at the
 * local ip; JTInv at the local ip.  Called from prep_elem, fv1/navier_stokes_fv1.cpp:208-248.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int elem, dim, nsh, nip;
    double x[MAXSH][3];
    int from[MAXIP], to[MAXIP];
    double n[MAXIP][3], xip[MAXIP][3], N[MAXIP][MAXSH], G[MAXIP][MAXSH][3], c0c2sq[MAXIP];
    double vol[MAXSH];
} Geom;

static int geom_update(Geom *g, int elem, const double *coords)
{
    const RefElem *r = get_ref(elem);
    if (!r) return fail("geom_update: unknown element type");
    g->elem = elem; g->dim = r->dim; g->nsh = r->nsh; g->nip = r->nedge;
    int dim = r->dim, nsh = r->nsh;
    for (int i = 0; i < nsh; i++) { for (int d = 0; d < dim; d++) g->x[i][d] = coords[i*dim+d]; for (int d = dim; d < 3; d++) g->x[i][d] = 0; }
    for (int ip = 0; ip < g->nip; ip++) {
        double c[4][3];
        int nc = scvf_corners(r, ip, g->x, c);
        g->from[ip] = r->edge[ip][0]; g->to[ip] = r->edge[ip][1];
        for (int d = 0; d < 3; d++) { double s = 0; for (int k = 0; k < nc; k++) s += c[k][d]; g->xip[ip][d] = s / nc; }
        scvf_normal(dim, c, g->n[ip]);
        g->c0c2sq[ip] = dim == 3 ? vdistsq(c[0], c[2], 3) : 0.0;
        /* JT(i,j) = sum_k dN_k/dxi_i * x_k[j]; global_grad = JT^{-1} * local_grad */
        double JT[3][3] = {{0}}, JTinv[3][3] = {{0}};
        for (int i = 0; i < dim; i++) for (int j = 0; j < dim; j++) {
            double s = 0; for (int k = 0; k < nsh; k++) s += r->lgrad_ip[ip][k][i] * g->x[k][j];
            JT[i][j] = s;
        }
        double det = mat_inverse(dim, JT, JTinv);
        if (!(fabs(det) > 0)) return fail("FV1Geometry: singular element Jacobian");
        for (int k = 0; k < nsh; k++) {
            g->N[ip][k] = r->shape_ip[ip][k];
            for (int j = 0; j < 3; j++) g->G[ip][k][j] = 0;
            for (int j = 0; j < dim; j++) { double s = 0; for (int i = 0; i < dim; i++) s += JTinv[j][i]*r->lgrad_ip[ip][k][i]; g->G[ip][k][j] = s; }
        }
    }
    /* SCV volumes */
    if (elem == ORA_TRI) {
        double a = 0.5*fabs((g->x[1][0]-g->x[0][0])*(g->x[2][1]-g->x[0][1]) - (g->x[2][0]-g->x[0][0])*(g->x[1][1]-g->x[0][1]));
        for (int i = 0; i < 3; i++) g->vol[i] = a/3.0;
    } else if (elem == ORA_TET) {
        double a[3], b[3], c[3], t[3];
        for (int d=0; d<3; d++) { a[d]=g->x[1][d]-g->x[0][d]; b[d]=g->x[2][d]-g->x[0][d]; c[d]=g->x[3][d]-g->x[0][d]; }
        vcross(t, a, b);
        double v = fabs(vdot(t, c, 3))/6.0;
        for (int i = 0; i < 4; i++) g->vol[i] = v/4.0;
    } else if (elem == ORA_QUAD) {
        /* SCV = quadrilateral (corner, mid of outgoing edge, barycentre, mid of incoming edge) */
        double bc[3]; int all[4] = {0,1,2,3}; avg_pts(bc, g->x, all, 4, 2);
        for (int i = 0; i < 4; i++) {
            int e_out[2] = {i, (i+1)%4}, e_in[2] = {(i+3)%4, i};
            double m1[3], m2[3]; avg_pts(m1, g->x, e_out, 2, 2); avg_pts(m2, g->x, e_in, 2, 2);
            /* 0.5*|(c2-c0) x (c3-c1)| with c = (corner, m1, bary, m2) */
            double ax = bc[0]-g->x[i][0], ay = bc[1]-g->x[i][1], bx = m2[0]-m1[0], by = m2[1]-m1[1];
            g->vol[i] = 0.5*fabs(ax*by - ay*bx);
        }
    } else if (elem == ORA_PRISM) {
        /* SCV of corner i = hexahedron (corner, edge midpoint, triangle centre, edge midpoint | axis-edge midpoint, quadrilateral
           centre, barycentre, quadrilateral centre); volume of the trilinear hexahedron through these eight points */
        int all[6] = {0,1,2,3,4,5}; double bc[3]; avg_pts(bc, g->x, all, 6, 3);
        for (int i = 0; i < 6; i++) {
            const int base = i < 3 ? 0 : 3, t = i - base, a = base + (t+1)%3, b = base + (t+2)%3, up = i < 3 ? i+3 : i-3;
            const int tri = i < 3 ? 0 : 4;
            int qa = -1, qb = -1;
            for (int s2 = 1; s2 <= 3; s2++) {
                int hi = 0, ha = 0, hb = 0;
                for (int k = 0; k < 4; k++) { hi |= r->side[s2][k]==i; ha |= r->side[s2][k]==a; hb |= r->side[s2][k]==b; }
                if (hi && ha) qa = s2;
                if (hi && hb) qb = s2;
            }
            double p[8][3]; int e2[2];
            for (int d = 0; d < 3; d++) p[0][d] = g->x[i][d];
            e2[0] = i; e2[1] = a;  avg_pts(p[1], g->x, e2, 2, 3);
            avg_pts(p[2], g->x, r->side[tri], 3, 3);
            e2[1] = b;             avg_pts(p[3], g->x, e2, 2, 3);
            e2[1] = up;            avg_pts(p[4], g->x, e2, 2, 3);
            avg_pts(p[5], g->x, r->side[qa], 4, 3);
            for (int d = 0; d < 3; d++) p[6][d] = bc[d];
            avg_pts(p[7], g->x, r->side[qb], 4, 3);
            g->vol[i] = fabs(hex_volume(p));
        }
    } else {
        /* SCV of corner i = trilinear image of the reference octant adjacent to corner i */
        for (int i = 0; i < 8; i++) {
            double p[8][3];
            for (int q = 0; q < 8; q++) {
                double xi[3], N[8];
                for (int d = 0; d < 3; d++) {
                    double lo = r->corner[i][d] < 0.5 ? 0.0 : 0.5, hi = lo + 0.5;
                    xi[d] = HEX_CO[q][d] > 0.5 ? hi : lo;
                }
                lagrange_shapes(ORA_HEX, xi, N, NULL);
                for (int d = 0; d < 3; d++) { double s = 0; for (int k = 0; k < 8; k++) s += N[k]*g->x[k][d]; p[q][d] = s; }
            }
            g->vol[i] = fabs(hex_volume(p));
        }
    }
    return 0;
}

int ora_fv1_geometry(int elem, const double *coords, ora_fv1_geom *out)
{
    Geom g; const RefElem *r = get_ref(elem);
    if (geom_update(&g, elem, coords)) return -1;
    memset(out, 0, sizeof *out);
    out->dim = g.dim; out->nsh = g.nsh; out->nip = g.nip;
    for (int ip = 0; ip < g.nip; ip++) {
        out->from[ip] = g.from[ip]; out->to[ip] = g.to[ip]; out->c0c2sq[ip] = g.c0c2sq[ip];
        for (int d = 0; d < 3; d++) { out->normal[ip][d] = g.n[ip][d]; out->xip[ip][d] = g.xip[ip][d]; out->lip[ip][d] = r->lip[ip][d]; }
        for (int k = 0; k < g.nsh; k++) { out->shape[ip][k] = g.N[ip][k]; for (int d = 0; d < 3; d++) out->ggrad[ip][k][d] = g.G[ip][k][d]; }
    }
    for (int k = 0; k < g.nsh; k++) out->vol[k] = g.vol[k];
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Boundary faces of FV1Geometry (ugcore fv1_geom.cpp, BF -- absent here, OUR SPEC like App. B-2):
 * a boundary side contributes one BF per side corner. 2-D: segment [corner, edge midpoint];
 * 3-D: quadrilateral [corner, midpoint of the edge to the next side corner, side centre, midpoint
 * of the edge to the previous side corner]. ip = mean of the BF corners (same construction on the
 * reference element for the local ip); normal = (dy,-dx) / 0.5 (c2-c0)x(c3-c1), turned so that it
 * points away from the element barycentre; volume = |normal|; shapes and global gradients of ALL
 * element shape functions at the BF ip.
 * Used by NavierStokesNoNormalStressOutflowFV1 (fv1/bnd/no_normal_stress_outflow_fv1.cpp:192-427)
 * and the NeumannBoundaryFV1 part of NavierStokesInflowFV1 (fv1/bnd/inflow_fv1_impl.h:42-82).
 * ---------------------------------------------------------------------------------------- */
typedef struct { int node_id; double n[3], xip[3], lip[3], N[MAXSH], G[MAXSH][3], vol; } BFace;
static inline int64_t csr_find(const int64_t *rowptr, const int32_t *colind, int64_t row, int32_t col);
static double mat_inverse(int dim, const double a[3][3], double inv[3][3]);

static void bf_corners(const RefElem *r, int side, int j, const double (*x)[3], double (*c)[3], int *nc)
{
    const int ns = r->side_n[side], co = r->side[side][j];
    for (int d = 0; d < 3; d++) c[0][d] = x[co][d];
    if (r->dim == 2) {
        const int other = r->side[side][1 - j];
        for (int d = 0; d < 3; d++) c[1][d] = 0.5 * (x[co][d] + x[other][d]);
        *nc = 2; return;
    }
    const int nx = r->side[side][(j + 1) % ns], pv = r->side[side][(j + ns - 1) % ns];
    for (int d = 0; d < 3; d++) { c[1][d] = 0.5 * (x[co][d] + x[nx][d]); c[3][d] = 0.5 * (x[co][d] + x[pv][d]); }
    avg_pts(c[2], x, r->side[side], ns, 3);
    *nc = 4;
}

static int bf_update(BFace *bf, int elem, const double *coords, int side, int j)
{
    const RefElem *r = get_ref(elem);
    if (!r) return fail("bf_update: unknown element type");
    if (side < 0 || side >= r->nside || j < 0 || j >= r->side_n[side]) return fail("bf_update: bad side / corner");
    const int dim = r->dim, nsh = r->nsh;
    double x[MAXSH][3], c[4][3], lc[4][3], bary[3], sc[3];
    int nc, all[MAXSH];
    for (int i = 0; i < nsh; i++) { all[i] = i; for (int d = 0; d < 3; d++) x[i][d] = d < dim ? coords[i*dim+d] : 0.0; }
    bf->node_id = r->side[side][j];
    bf_corners(r, side, j, x, c, &nc);
    bf_corners(r, side, j, r->corner, lc, &nc);
    for (int d = 0; d < 3; d++) { double s = 0, t = 0; for (int k = 0; k < nc; k++) { s += c[k][d]; t += lc[k][d]; } bf->xip[d] = s / nc; bf->lip[d] = t / nc; }
    scvf_normal(dim, c, bf->n);
    avg_pts(bary, x, all, nsh, dim); avg_pts(sc, x, r->side[side], r->side_n[side], dim);
    { double o[3] = {sc[0]-bary[0], sc[1]-bary[1], dim == 3 ? sc[2]-bary[2] : 0.0};
      if (vdot(bf->n, o, dim) < 0) for (int d = 0; d < 3; d++) bf->n[d] = -bf->n[d]; }
    bf->vol = sqrt(vdot(bf->n, bf->n, dim));
    double lg[MAXSH][3], JT[3][3] = {{0}}, JTinv[3][3] = {{0}};
    lagrange_shapes(elem, bf->lip, bf->N, lg);
    for (int i = 0; i < dim; i++) for (int jj = 0; jj < dim; jj++) { double s = 0; for (int k = 0; k < nsh; k++) s += lg[k][i] * x[k][jj]; JT[i][jj] = s; }
    if (!(fabs(mat_inverse(dim, JT, JTinv)) > 0)) return fail("FV1Geometry: singular element Jacobian");
    for (int k = 0; k < nsh; k++) for (int jj = 0; jj < 3; jj++) {
        double s = 0; if (jj < dim) for (int i = 0; i < dim; i++) s += JTinv[jj][i] * lg[k][i];
        bf->G[k][jj] = s;
    }
    return 0;
}

int ora_fv1_bf_geometry(int elem, const double *coords, int side, int j, int *node_id, double *normal, double *xip,
                        double *shape, double *ggrad)
{
    BFace bf;
    if (bf_update(&bf, elem, coords, side, j)) return -1;
    const int dim = ora_elem_dim(elem), nsh = ora_elem_nsh(elem);
    *node_id = bf.node_id;
    for (int d = 0; d < dim; d++) { normal[d] = bf.n[d]; xip[d] = bf.xip[d]; }
    for (int k = 0; k < nsh; k++) { shape[k] = bf.N[k]; for (int d = 0; d < dim; d++) ggrad[k*dim+d] = bf.G[k][d]; }
    return 0;
}

int ora_side_corners(int elem, int side)
{ const RefElem *r = get_ref(elem); return (r && side >= 0 && side < r->nside) ? r->side_n[side] : -1; }
int ora_side_corner(int elem, int side, int j)
{ const RefElem *r = get_ref(elem); return (r && side >= 0 && side < r->nside && j >= 0 && j < r->side_n[side]) ? r->side[side][j] : -1; }

/* Boundary contributions of the listed (element, side) pairs, ADDED to values / defect (scaled by scale_a).
 * kind 0: NavierStokesNoNormalStressOutflowFV1::add_jac_A_elem / add_def_A_elem
 *         (fv1/bnd/no_normal_stress_outflow_fv1.cpp:343-427 with diffusive_flux_Jac :192-236, diffusive_flux_defect :239-279,
 *          convective_flux_Jac :282-313, convective_flux_defect :316-338); constant viscosity / density.
 * kind 1: NeumannBoundaryFV1 with vector data on the pressure function (the continuity-equation part of NavierStokesInflowFV1,
 *         fv1/bnd/inflow_fv1_impl.h:42-82; ugcore neumann_boundary_fv1.cpp VectorData::add_rhs_elem, absent: rhs(p, co) -= data . n,
 *         i.e. defect(p, co) += scale_a data . n). data [n_side][4][dim] at the BF ips (side-corner order, unused slots ignored). */
int ora_fv1_boundary(const ora_params *p, int kind, int64_t n_side, const int32_t *belem, const int32_t *bside, const double *data,
                     const int32_t *conn, const double *coords, const double *u, const int64_t *rowptr, const int32_t *colind,
                     int what, double scale_a, double *values, double *defect)
{
    const RefElem *r = get_ref(p->elem);
    if (!r) return fail("ora_fv1_boundary: unknown element type");
    const int dim = r->dim, nsh = r->nsh, nf = dim + 1, P = dim;
    for (int64_t b = 0; b < n_side; b++) {
        const int64_t e = belem[b]; const int side = bside[b];
        double xc[MAXSH*3], ul[MAXSH][4];
        for (int k = 0; k < nsh; k++) {
            const int64_t nd = conn[e*nsh+k];
            for (int d = 0; d < dim; d++) xc[k*dim+d] = coords[nd*dim+d];
            for (int f = 0; f < nf; f++) ul[k][f] = u ? u[nd*nf+f] : 0.0;
        }
        if (side < 0 || side >= r->nside) return fail("ora_fv1_boundary: bad side");
        for (int j = 0; j < r->side_n[side]; j++) {
            BFace bf;
            if (bf_update(&bf, p->elem, xc, side, j)) return -1;
            const int co = bf.node_id; const int64_t row0 = (int64_t)conn[e*nsh+co]*nf;
            if (kind == 1) {
                if (what & (ORA_DEF_A|ORA_RHS)) { double s = 0; for (int d = 0; d < dim; d++) s += data[(b*4+j)*dim+d] * bf.n[d]; defect[row0+P] += scale_a * s; }
                continue;
            }
            const double nurho = p->kin_visc * p->density;
            double std_[3] = {0,0,0};
            for (int k = 0; k < nsh; k++) for (int d = 0; d < dim; d++) std_[d] += ul[k][d] * bf.N[k];
            double flux = vdot(std_, bf.n, dim) * p->density;                  /* :296, :326 */
            if (what & ORA_JAC_A) {
                for (int sh = 0; sh < nsh; sh++) {
                    double T[3][3], ns_[3];
                    const double gn = vdot(bf.G[sh], bf.n, dim);
                    for (int d1 = 0; d1 < dim; d1++) for (int d2 = 0; d2 < dim; d2++) {
                        T[d1][d2] = d1 == d2 ? gn : 0.0;
                        if (!p->laplace) T[d1][d2] += bf.G[sh][d1] * bf.n[d2];
                    }
                    for (int d2 = 0; d2 < dim; d2++) { double s = 0; for (int d1 = 0; d1 < dim; d1++) s += T[d1][d2] * bf.n[d1]; ns_[d2] = s; }   /* TransposedMatVecMult :219 */
                    const int64_t col0 = (int64_t)conn[e*nsh+sh]*nf;
                    for (int d1 = 0; d1 < dim; d1++) for (int d2 = 0; d2 < dim; d2++) {
                        double v = (T[d1][d2] - bf.n[d1] * ns_[d2]) * (-nurho);
                        if (d1 == d2 && !p->stokes) v += (flux < 0 ? 0.0 : flux) * bf.N[sh];
                        int64_t q = csr_find(rowptr, colind, row0+d1, (int32_t)(col0+d2));
                        if (q < 0) return fail("ora_fv1_boundary: entry not in CSR pattern");
                        values[q] += scale_a * v;
                    }
                    for (int d2 = 0; d2 < dim; d2++) {
                        int64_t q = csr_find(rowptr, colind, row0+P, (int32_t)(col0+d2));
                        if (q < 0) return fail("ora_fv1_boundary: entry not in CSR pattern");
                        values[q] += scale_a * bf.N[sh] * bf.n[d2] * p->density;
                    }
                }
            }
            if (what & ORA_DEF_A) {
                double gv[3][3], df[3];
                for (int d1 = 0; d1 < dim; d1++) for (int d2 = 0; d2 < dim; d2++) { double s = 0; for (int sh = 0; sh < nsh; sh++) s += bf.G[sh][d2] * ul[sh][d1]; gv[d1][d2] = s; }
                for (int d1 = 0; d1 < dim; d1++) {
                    double s = 0; for (int d2 = 0; d2 < dim; d2++) s += gv[d1][d2] * bf.n[d2];
                    if (!p->laplace) for (int d2 = 0; d2 < dim; d2++) s += gv[d2][d1] * bf.n[d2];
                    df[d1] = s;
                }
                const double dn = vdot(df, bf.n, dim);
                for (int d1 = 0; d1 < dim; d1++) {
                    double v = (df[d1] - dn * bf.n[d1]) * (-nurho);                  /* VecScaleAppend(diffFlux, -dot, normal) :270: NOT normalised */
                    if (!p->stokes) v += (flux < 0 ? 0.0 : flux) * std_[d1];
                    defect[row0+d1] += scale_a * v;
                }
                defect[row0+P] += scale_a * vdot(std_, bf.n, dim) * p->density;
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * ElementSideRayIntersection (ugcore lib_disc/common/geometry_util.h -- App. B-4, our spec).
 * Sides are visited in reference order; 3-D sides are tested as triangle (p0,p1,p2) and, for
 * quadrilateral sides, (p0,p2,p3); the first hit with t<=0 (upwind search, bPositive=false)
 * or t>=0 wins. local cut = barycentric combination of the reference corners.
 * Called at upwind.cpp:351,420,470,547,615.
 * ---------------------------------------------------------------------------------------- */
#define RAY_SMALL 1e-12
static int ray_line_2d(const double *p0, const double *p1, const double *from, const double *dir,
                       double *bc, double *t)
{
    /* from + t*dir = p0 + bc*(p1-p0) */
    double ex = p1[0]-p0[0], ey = p1[1]-p0[1];
    double det = dir[0]*(-ey) - dir[1]*(-ex);         /* | dir  -e | */
    double scale = sqrt(dir[0]*dir[0]+dir[1]*dir[1]) * sqrt(ex*ex+ey*ey);
    if (!(fabs(det) > RAY_SMALL*scale)) return 0;
    double rx = p0[0]-from[0], ry = p0[1]-from[1];
    *t  = (rx*(-ey) - ry*(-ex)) / det;
    *bc = (dir[0]*ry - dir[1]*rx) / det;
    return (*bc >= -RAY_SMALL && *bc <= 1.0 + RAY_SMALL);
}
static int ray_triangle(const double *p0, const double *p1, const double *p2, const double *from,
                        const double *dir, double *b1, double *b2, double *t)
{
    /* from + t*dir = p0 + b1*e1 + b2*e2  (Cramer) */
    double e1[3], e2[3], r[3], nrm[3], q[3];
    for (int d = 0; d < 3; d++) { e1[d]=p1[d]-p0[d]; e2[d]=p2[d]-p0[d]; r[d]=from[d]-p0[d]; }
    vcross(nrm, e1, e2);
    double det = -vdot(dir, nrm, 3);
    double scale = sqrt(vdot(dir,dir,3)) * sqrt(vdot(nrm,nrm,3));
    if (!(fabs(det) > RAY_SMALL*scale)) return 0;
    *t = vdot(r, nrm, 3) / det;
    vcross(q, r, dir);                    /* q = r x dir */
    *b1 =  vdot(e2, q, 3) / det;
    *b2 = -vdot(e1, q, 3) / det;
    return (*b1 >= -RAY_SMALL && *b2 >= -RAY_SMALL && *b1 + *b2 <= 1.0 + RAY_SMALL);
}

static int side_ray_intersection(const RefElem *r, const double (*x)[3], const double *from,
                                 const double *dir, int positive, int *side_out, double *gcut, double *lcut)
{
    int dim = r->dim;
    for (int s = 0; s < r->nside; s++) {
        if (dim == 2) {
            int p0 = r->side[s][0], p1 = r->side[s][1]; double bc, t;
            if (!ray_line_2d(x[p0], x[p1], from, dir, &bc, &t)) continue;
            if (!((t >= 0.0 && positive) || (t <= 0.0 && !positive))) continue;
            for (int d = 0; d < 2; d++) { gcut[d] = from[d] + t*dir[d]; lcut[d] = (1-bc)*r->corner[p0][d] + bc*r->corner[p1][d]; }
            gcut[2] = lcut[2] = 0; *side_out = s; return 1;
        } else {
            int ntri = r->side_n[s] == 4 ? 2 : 1;
            for (int k = 0; k < ntri; k++) {
                int p0 = r->side[s][0], p1 = r->side[s][1+k], p2 = r->side[s][2+k]; double b1, b2, t;
                if (!ray_triangle(x[p0], x[p1], x[p2], from, dir, &b1, &b2, &t)) continue;
                if (!((t >= 0.0 && positive) || (t <= 0.0 && !positive))) continue;
                for (int d = 0; d < 3; d++) { gcut[d] = from[d] + t*dir[d];
                    lcut[d] = (1-b1-b2)*r->corner[p0][d] + b1*r->corner[p1][d] + b2*r->corner[p2][d]; }
                *side_out = s; return 1;
            }
        }
    }
    return 0;
}

int ora_side_ray_intersection(int elem, const double *coords, const double *from, const double *dir,
                              int positive, int *side, double *gcut, double *lcut)
{
    const RefElem *r = get_ref(elem); if (!r) return -1;
    double x[MAXSH][3] = {{0}}, f[3] = {0}, dr[3] = {0};
    for (int i = 0; i < r->nsh; i++) for (int d = 0; d < r->dim; d++) x[i][d] = coords[i*r->dim+d];
    for (int d = 0; d < r->dim; d++) { f[d] = from[d]; dr[d] = dir[d]; }
    double g[3], l[3];
    int ok = side_ray_intersection(r, x, f, dr, positive, side, g, l);
    for (int d = 0; d < r->dim; d++) { gcut[d] = g[d]; lcut[d] = l[d]; }
    return ok;
}

/* ------------------------------------------------------------------------------------------
 * Upwinds on FV1 geometry (upwind.cpp)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    double sh[MAXIP][MAXSH];
    double ip[MAXIP][MAXIP];
    double len[MAXIP];
    int nonzero_ip;                 /* INavierStokesUpwind::non_zero_shape_ip() */
} Upw;

/* upwind.cpp:52-80 */
static void upwind_no(const Geom *g, Upw *u)
{
    for (int ip = 0; ip < g->nip; ip++) { for (int sh = 0; sh < g->nsh; sh++) u->sh[ip][sh] = g->N[ip][sh]; u->len[ip] = 1.0; }
}
/* upwind.cpp:133-172 */
static void upwind_full(const Geom *g, const double (*vel)[3], Upw *u)
{
    for (int ip = 0; ip < g->nip; ip++) {
        for (int sh = 0; sh < g->nsh; sh++) u->sh[ip][sh] = 0.0;
        double flux = vdot(g->n[ip], vel[ip], g->dim);
        int co = flux > 0.0 ? g->from[ip] : g->to[ip];
        u->sh[ip][co] = 1.0;
        u->len[ip] = vdist(g->xip[ip], g->x[co], g->dim);
    }
}
/* upwind.cpp:337-430 (GetNodeNextToCut + NavierStokesSkewedUpwind::compute) */
static int upwind_skewed(const Geom *g, const double (*vel)[3], Upw *u)
{
    const RefElem *r = get_ref(g->elem);
    for (int ip = 0; ip < g->nip; ip++) {
        for (int sh = 0; sh < g->nsh; sh++) u->sh[ip][sh] = 0.0;
        if (sqrt(vdot(vel[ip], vel[ip], g->dim)) < 1e-14) { u->len[ip] = 1.0; continue; }
        int side; double gc[3], lc[3];
        if (!side_ray_intersection(r, g->x, g->xip[ip], vel[ip], 0, &side, gc, lc))
            return fail("GetNodeNextToCut: Cannot find cut side.");
        double min = DBL_MAX; int co_out = 0;
        for (int i = 0; i < r->side_n[side]; i++) {
            int co = r->side[side][i];
            double dist = vdistsq(gc, g->x[co], g->dim);
            if (dist < min) { min = dist; co_out = co; }
        }
        u->sh[ip][co_out] = 1.0;
        u->len[ip] = vdist(g->xip[ip], g->x[co_out], g->dim);
    }
    return 0;
}
/* upwind.cpp:505-575 */
static int upwind_lps(const Geom *g, const double (*vel)[3], Upw *u)
{
    const RefElem *r = get_ref(g->elem);
    for (int ip = 0; ip < g->nip; ip++) {
        for (int sh = 0; sh < g->nsh; sh++) u->sh[ip][sh] = 0.0;
        if (sqrt(vdot(vel[ip], vel[ip], g->dim)) < 1e-14) { u->len[ip] = 1.0; continue; }
        int side; double gc[3], lc[3], N[MAXSH];
        if (!side_ray_intersection(r, g->x, g->xip[ip], vel[ip], 0, &side, gc, lc))
            return fail("GetLinearProfileSkewedUpwindShapes: Cannot find cut side.");
        lagrange_shapes(g->elem, lc, N, NULL);
        for (int j = 0; j < r->side_n[side]; j++) { int co = r->side[side][j]; u->sh[ip][co] = N[co]; }
        u->len[ip] = vdist(g->xip[ip], gc, g->dim);
    }
    return 0;
}
/* upwind.cpp:643-786 */
static void upwind_positive(const Geom *g, const double (*vel)[3], Upw *u)
{
    double flux[MAXIP]; int has[MAXIP]; int n_noflux = 0;
    const double eps = DBL_EPSILON * 10;
    int nip = g->nip, nsh = g->nsh, dim = g->dim;
    for (int ip = 0; ip < nip; ip++) {
        flux[ip] = 0.0; has[ip] = 1;
        for (int sh = 0; sh < nsh; sh++) u->sh[ip][sh] = 0.0;
        for (int j = 0; j < nip; j++) u->ip[ip][j] = 0.0;
        double normsq = vdot(vel[ip], vel[ip], dim);
        if (fabs(normsq) <= eps) { u->sh[ip][g->from[ip]] = 0.5; u->sh[ip][g->to[ip]] = 0.5; has[ip] = 0; n_noflux++; continue; }
        flux[ip] = vdot(vel[ip], g->n[ip], dim);
        double v = sqrt(normsq), len = sqrt(vdot(g->n[ip], g->n[ip], dim));
        if (fabs(flux[ip] / sqrt(v*len)) <= eps) { u->sh[ip][g->from[ip]] = 0.5; u->sh[ip][g->to[ip]] = 0.5; has[ip] = 0; n_noflux++; continue; }
    }
    if (n_noflux != nip) {
        for (int sh = 0; sh < nsh; sh++) {
            double m_in = 0, m_out = 0; int ips[MAXIP]; double fl[MAXIP]; int cnt = 0;
            for (int ip = 0; ip < nip; ip++) {
                if (!has[ip]) continue;
                if (g->from[ip] == sh) { ips[cnt] = ip; fl[cnt++] = flux[ip]; m_in += -1.0*fmin(flux[ip], 0.0); m_out += fmax(flux[ip], 0.0); }
                else if (g->to[ip] == sh) { ips[cnt] = ip; fl[cnt++] = -1.0*flux[ip]; m_in += -1.0*fmin(-1.0*flux[ip], 0.0); m_out += fmax(-1.0*flux[ip], 0.0); }
            }
            double F = fmax(m_in, m_out);
            for (int i = 0; i < cnt; i++) if (fl[i] > 0) {
                double sum = 0.0;
                for (int j = 0; j < cnt; j++) if (fl[j] < 0) { u->ip[ips[i]][ips[j]] = -1.0*fl[j]/F; sum += u->ip[ips[i]][ips[j]]; }
                u->sh[ips[i]][sh] = 1.0 - sum;
            }
        }
    }
    for (int ip = 0; ip < nip; ip++) {
        double up[3] = {0,0,0};
        for (int sh = 0; sh < nsh; sh++) for (int d = 0; d < dim; d++) up[d] += u->sh[ip][sh]*g->x[sh][d];
        for (int j = 0; j < nip; j++) for (int d = 0; d < dim; d++) up[d] += u->ip[ip][j]*g->xip[j][d];
        u->len[ip] = vdist(g->xip[ip], up, dim);
    }
}

static int upwind_compute(int type, const Geom *g, const double (*vel)[3], Upw *u)
{
    /* non-Positive upwinds never write the ip shapes (upwind.h:69,114,196,232): poison them */
    u->nonzero_ip = (type == ORA_UPWIND_POSITIVE);
    if (!u->nonzero_ip) for (int i = 0; i < MAXIP; i++) for (int j = 0; j < MAXIP; j++) u->ip[i][j] = NAN;
    switch (type) {
    case ORA_UPWIND_NO: upwind_no(g, u); return 0;
    case ORA_UPWIND_FULL: upwind_full(g, vel, u); return 0;
    case ORA_UPWIND_SKEWED: return upwind_skewed(g, vel, u);
    case ORA_UPWIND_LPS: return upwind_lps(g, vel, u);
    case ORA_UPWIND_POSITIVE: upwind_positive(g, vel, u); return 0;
    }
    return fail("upwind type unknown / not set");
}

/* INavierStokesUpwind::upwind_vel, upwind_interface.h:334-358 */
static void upwind_vel(const Upw *uw, const Geom *g, int ip, const double *u /*[fct][sh]*/,
                       const double (*stdvel)[3], double *vel)
{
    for (int d = 0; d < 3; d++) vel[d] = 0;
    for (int sh = 0; sh < g->nsh; sh++) for (int d = 0; d < g->dim; d++) vel[d] += uw->sh[ip][sh]*u[d*g->nsh+sh];
    if (!uw->nonzero_ip) return;
    for (int j = 0; j < g->nip; j++) for (int d = 0; d < g->dim; d++) vel[d] += uw->ip[ip][j]*stdvel[j][d];
}

int ora_fv1_upwind(int elem, int upwind, const double *coords, const double *ipvel,
                   double *up_sh, double *up_ip, double *conv_len)
{
    Geom g; if (geom_update(&g, elem, coords)) return -1;
    double vel[MAXIP][3] = {{0}};
    for (int ip = 0; ip < g.nip; ip++) for (int d = 0; d < g.dim; d++) vel[ip][d] = ipvel[ip*g.dim+d];
    Upw u; if (upwind_compute(upwind, &g, vel, &u)) return -1;
    for (int ip = 0; ip < g.nip; ip++) {
        for (int sh = 0; sh < g.nsh; sh++) up_sh[ip*g.nsh+sh] = u.sh[ip][sh];
        for (int j = 0; j < g.nip; j++) up_ip[ip*g.nip+j] = u.nonzero_ip ? u.ip[ip][j] : 0.0;
        conv_len[ip] = u.len[ip];
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Diffusion lengths (fv1/diffusion_length.h:47-198)
 * ---------------------------------------------------------------------------------------- */
static int diff_length(int type, const Geom *g, double *out)
{
    int dim = g->dim, nip = g->nip;
    double minN = DBL_MAX, minD = DBL_MAX, avgN = 0.0;
    if (type == ORA_DIFF_COR) {
        for (int i = 0; i < nip; i++) {
            double nn = vdot(g->n[i], g->n[i], dim);
            if (nn < minN) minN = nn;
            avgN += nn;
            if (dim == 3 && g->c0c2sq[i] < minD) minD = g->c0c2sq[i];
        }
        avgN /= nip;
    }
    for (int i = 0; i < nip; i++) {
        double nn = vdot(g->n[i], g->n[i], dim);
        double a = 0.5*(g->vol[g->from[i]] + g->vol[g->to[i]]); a *= a;
        double ds = g->c0c2sq[i];
        switch (type) {
        case ORA_DIFF_FIVEPOINT: out[i] = dim == 2 ? 2.0*nn/a + 8.0/nn : 2.0*nn/a + 8.0*ds/nn; break;
        case ORA_DIFF_RAW:       out[i] = dim == 2 ? 1./(0.5*a/nn + 3.0*nn/8.) : 1./(0.5*a/nn + 3.0*ds/8.); break;
        case ORA_DIFF_COR:       out[i] = dim == 2 ? 2.0*minN/a + 8.0/(3.0*avgN) : 2.0*minN/a + 8.0*minD/(3.0*avgN); break;
        default: return fail(" Diffusion Length type not found.");
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Dense inverse (ugcore GetInverse/MatMult, App. B-5): LU with partial pivoting
 * ---------------------------------------------------------------------------------------- */
typedef struct { int n; double lu[MAXIP][MAXIP]; int piv[MAXIP]; } LU;
static int lu_factor(LU *f, int n, double (*m)[MAXIP])
{
    f->n = n;
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) f->lu[i][j] = m[i][j];
    for (int k = 0; k < n; k++) {
        int p = k; double best = fabs(f->lu[k][k]);
        for (int i = k+1; i < n; i++) if (fabs(f->lu[i][k]) > best) { best = fabs(f->lu[i][k]); p = i; }
        if (!(best > 0.0)) return -1;
        f->piv[k] = p;
        if (p != k) for (int j = 0; j < n; j++) { double t = f->lu[k][j]; f->lu[k][j] = f->lu[p][j]; f->lu[p][j] = t; }
        for (int i = k+1; i < n; i++) {
            f->lu[i][k] /= f->lu[k][k];
            for (int j = k+1; j < n; j++) f->lu[i][j] -= f->lu[i][k]*f->lu[k][j];
        }
    }
    return 0;
}
static void lu_solve(const LU *f, const double *b, double *x)
{
    int n = f->n;
    for (int i = 0; i < n; i++) x[i] = b[i];
    for (int k = 0; k < n; k++) { int p = f->piv[k]; if (p != k) { double t = x[k]; x[k] = x[p]; x[p] = t; } }
    for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) x[i] -= f->lu[i][j]*x[j];
    for (int i = n-1; i >= 0; i--) { for (int j = i+1; j < n; j++) x[i] -= f->lu[i][j]*x[j]; x[i] /= f->lu[i][i]; }
}

/* ------------------------------------------------------------------------------------------
 * FV1 stabilisations (fv1/stabilization.cpp)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    double vel[MAXIP][3];
    double sv[MAXIP][3][3][MAXSH];     /* stab_shape_vel(ip, compOut, compIn, sh) */
    double sp[MAXIP][3][MAXSH];        /* stab_shape_p(ip, compOut, sh) */
    int connected;                     /* vel_comp_connected() */
    Upw up, down;                      /* the stabilisation's upwind object state */
} Stab;

#define U_(f,sh)  (u[(f)*nsh+(sh)])
/* per-ip data imports (m_imKinViscosity[ip], m_imDensitySCVF[ip], m_imDensitySCV[ip], m_imSourceSCV(F)[ip]) or the constants */
#define VISC(ip)    (p->ip_visc ? p->ip_visc[p->elem_index*nip+(ip)] : p->kin_visc)
#define RHOF(ip)    (p->ip_rho_scvf ? p->ip_rho_scvf[p->elem_index*nip+(ip)] : p->density)
#define RHOV(sh)    (p->ip_rho_scv ? p->ip_rho_scv[p->elem_index*nsh+(sh)] : p->density)
#define HAS_SRCF    (p->has_source || p->ip_src_scvf)
#define SRCF(ip,d)  (p->ip_src_scvf ? p->ip_src_scvf[(p->elem_index*nip+(ip))*dim+(d)] : p->source[d])
#define HAS_SRCV    (p->has_source || p->ip_src_scv)
#define SRCV(sh,d)  (p->ip_src_scv ? p->ip_src_scv[(p->elem_index*nsh+(sh))*dim+(d)] : p->source[d])

/* NavierStokesFIELDSStabilization::update, stabilization.cpp:122-404 */
static int stab_fields(const ora_params *p, const Geom *g, const double *u /*vCornerValue*/,
                       const double (*stdvel)[3], int bStokes, const double *uold, double dt, Stab *s)
{
    int dim = g->dim, nsh = g->nsh, nip = g->nip, P = dim;
    s->connected = 0;
    if (!bStokes) if (upwind_compute(p->stab_upwind, g, stdvel, &s->up)) return -1;
    double dl[MAXIP]; if (diff_length(p->diff_len, g, dl)) return -1;
    double a[MAXIP], b[MAXIP];
    for (int ip = 0; ip < nip; ip++) a[ip] = VISC(ip) * dl[ip];
    if (!bStokes) for (int ip = 0; ip < nip; ip++) b[ip] = sqrt(vdot(stdvel[ip], stdvel[ip], dim)) / s->up.len[ip];
    /* off-diagonal vel shapes are never written by FIELDS (stale in the reference): poison */
    for (int ip = 0; ip < nip; ip++) for (int d = 0; d < 3; d++) for (int d2 = 0; d2 < 3; d2++)
        if (d != d2) for (int k = 0; k < MAXSH; k++) s->sv[ip][d][d2][k] = NAN;

    if (bStokes || !s->up.nonzero_ip) {
        for (int ip = 0; ip < nip; ip++) {
            double diag = a[ip];
            if (uold) diag += 1./dt;
            if (!bStokes) diag += b[ip];
            for (int d = 0; d < dim; d++) {
                double rhs = 0.0;
                if (HAS_SRCF) rhs = SRCF(ip,d);
                if (uold) { double o = 0.0; for (int sh = 0; sh < nsh; sh++) o += g->N[ip][sh]*uold[d*nsh+sh]; rhs += o/dt; }
                for (int k = 0; k < nsh; k++) {
                    double sumVel = a[ip]*g->N[ip][k];
                    if (!bStokes) sumVel += b[ip]*s->up.sh[ip][k];
                    rhs += sumVel*U_(d,k);
                    s->sv[ip][d][d][k] = sumVel/diag;
                    double sumP = -1.0*g->G[ip][k][d]/RHOF(ip);
                    rhs += sumP*U_(P,k);
                    s->sp[ip][d][k] = sumP/diag;
                }
                s->vel[ip][d] = rhs/diag;
            }
        }
    } else {
        double mat[MAXIP][MAXIP]; memset(mat, 0, sizeof mat);
        for (int ip = 0; ip < nip; ip++) {
            if (uold) mat[ip][ip] += 1./dt;
            mat[ip][ip] += a[ip];
            double scale = b[ip];
            mat[ip][ip] += scale;
            for (int ip2 = 0; ip2 < nip; ip2++) mat[ip][ip2] -= s->up.ip[ip][ip2]*scale;
        }
        LU lu; if (lu_factor(&lu, nip, mat)) return fail("Could not compute inverse.");
        for (int d = 0; d < dim; d++) {
            double cV[MAXSH][MAXIP], cP[MAXSH][MAXIP], x[MAXIP], f[MAXIP];
            for (int ip = 0; ip < nip; ip++) for (int k = 0; k < nsh; k++) {
                cV[k][ip] = a[ip]*g->N[ip][k];
                cV[k][ip] += b[ip]*s->up.sh[ip][k];
                cP[k][ip] = -1.0*g->G[ip][k][d]/RHOF(ip);
            }
            for (int k = 0; k < nsh; k++) {
                lu_solve(&lu, cV[k], x); for (int ip = 0; ip < nip; ip++) s->sv[ip][d][d][k] = x[ip];
                lu_solve(&lu, cP[k], x); for (int ip = 0; ip < nip; ip++) s->sp[ip][d][k] = x[ip];
            }
            for (int ip = 0; ip < nip; ip++) {
                f[ip] = 0.0;
                if (HAS_SRCF) f[ip] = SRCF(ip,d);
                if (uold) { double o = 0.0; for (int sh = 0; sh < nsh; sh++) o += g->N[ip][sh]*uold[d*nsh+sh]; f[ip] += o/dt; }
            }
            for (int k = 0; k < nsh; k++) for (int ip = 0; ip < nip; ip++) {
                f[ip] += U_(d,k)*cV[k][ip];
                f[ip] += U_(P,k)*cP[k][ip];
            }
            lu_solve(&lu, f, x);
            for (int ip = 0; ip < nip; ip++) s->vel[ip][d] = x[ip];
        }
    }
    return 0;
}

/* NavierStokesFLOWStabilization::update, stabilization.cpp:436-772 */
static int stab_flow(const ora_params *p, const Geom *g, const double *u, const double (*stdvel)[3],
                     int bStokes, const double *uold, double dt, Stab *s)
{
    int dim = g->dim, nsh = g->nsh, nip = g->nip, P = dim;
    s->connected = 1;
    if (!bStokes) {
        if (upwind_compute(p->stab_upwind, g, stdvel, &s->up)) return -1;
        double neg[MAXIP][3];                              /* update_downwind, upwind_interface.h:157-165 */
        for (int ip = 0; ip < nip; ip++) for (int d = 0; d < 3; d++) neg[ip][d] = -1.0*stdvel[ip][d];
        if (upwind_compute(p->stab_upwind, g, neg, &s->down)) return -1;
    }
    double dl[MAXIP]; if (diff_length(p->diff_len, g, dl)) return -1;
    double a[MAXIP], b[MAXIP], c[MAXIP];
    for (int ip = 0; ip < nip; ip++) a[ip] = VISC(ip) * dl[ip];
    if (!bStokes) for (int ip = 0; ip < nip; ip++) {
        double norm = sqrt(vdot(stdvel[ip], stdvel[ip], dim));
        b[ip] = norm / s->up.len[ip];
        c[ip] = norm / (s->down.len[ip] + s->up.len[ip]);
    }
    if (bStokes || !s->up.nonzero_ip) {
        for (int ip = 0; ip < nip; ip++) {
            double diag = a[ip];
            if (uold) diag += 1./dt;
            if (!bStokes) diag += b[ip];
            for (int d = 0; d < dim; d++) {
                double rhs = 0.0;
                if (HAS_SRCF) rhs = SRCF(ip,d);
                if (uold) { double o = 0.0; for (int sh = 0; sh < nsh; sh++) o += g->N[ip][sh]*uold[d*nsh+sh]; rhs += o/dt; }
                for (int k = 0; k < nsh; k++) {
                    double sumVel = a[ip]*g->N[ip][k];
                    if (!bStokes) {
                        sumVel += b[ip]*s->up.sh[ip][k];
                        sumVel += c[ip]*(s->down.sh[ip][k] - s->up.sh[ip][k]);
                    }
                    for (int d2 = 0; d2 < dim; d2++) { if (d2 == d) continue; sumVel -= stdvel[ip][d2]*g->G[ip][k][d2]; }
                    rhs += sumVel*U_(d,k);
                    s->sv[ip][d][d][k] = sumVel/diag;
                    for (int d2 = 0; d2 < dim; d2++) {
                        if (d2 == d) continue;
                        double sumVel2 = stdvel[ip][d]*g->G[ip][k][d2];
                        rhs += sumVel2*U_(d2,k);
                        s->sv[ip][d][d2][k] = sumVel2/diag;
                    }
                    double sumP = -1.0*g->G[ip][k][d]/RHOF(ip);
                    rhs += sumP*U_(P,k);
                    s->sp[ip][d][k] = sumP/diag;
                }
                s->vel[ip][d] = rhs/diag;
            }
        }
    } else {
        double mat[MAXIP][MAXIP]; memset(mat, 0, sizeof mat);
        for (int ip = 0; ip < nip; ip++) {
            if (uold) mat[ip][ip] += 1./dt;
            mat[ip][ip] += a[ip];
            mat[ip][ip] += b[ip];
            for (int ip2 = 0; ip2 < nip; ip2++) {
                mat[ip][ip2] -= s->up.ip[ip][ip2]*b[ip];
                mat[ip][ip2] += c[ip]*(s->up.ip[ip][ip2] - s->down.ip[ip][ip2]);
            }
        }
        LU lu; if (lu_factor(&lu, nip, mat)) return fail("Could not compute inverse.");
        double cV[3][MAXSH][MAXIP], cP[MAXSH][MAXIP], x[MAXIP], f[MAXIP];
        for (int d = 0; d < dim; d++) {
            for (int ip = 0; ip < nip; ip++) for (int k = 0; k < nsh; k++) {
                cP[k][ip] = -1.0*g->G[ip][k][d]/RHOF(ip);
                cV[d][k][ip] = a[ip]*g->N[ip][k];
                cV[d][k][ip] += b[ip]*s->up.sh[ip][k];
                cV[d][k][ip] += c[ip]*(s->down.sh[ip][k] - s->up.sh[ip][k]);
                for (int d2 = 0; d2 < dim; d2++) {
                    if (d2 == d) continue;
                    cV[d][k][ip] -= stdvel[ip][d2]*g->G[ip][k][d2];
                    cV[d2][k][ip] = stdvel[ip][d]*g->G[ip][k][d2];
                }
            }
            for (int k = 0; k < nsh; k++) {
                lu_solve(&lu, cP[k], x); for (int ip = 0; ip < nip; ip++) s->sp[ip][d][k] = x[ip];
                for (int d2 = 0; d2 < dim; d2++) { lu_solve(&lu, cV[d2][k], x); for (int ip = 0; ip < nip; ip++) s->sv[ip][d][d2][k] = x[ip]; }
            }
            for (int ip = 0; ip < nip; ip++) f[ip] = 0.0;
            for (int k = 0; k < nsh; k++) for (int ip = 0; ip < nip; ip++) {
                for (int d2 = 0; d2 < dim; d2++) f[ip] += U_(d2,k)*cV[d2][k][ip];
                f[ip] += U_(P,k)*cP[k][ip];
            }
            for (int ip = 0; ip < nip; ip++) {
                if (HAS_SRCF) f[ip] += SRCF(ip,d);
                if (uold) { double o = 0.0; for (int sh = 0; sh < nsh; sh++) o += g->N[ip][sh]*uold[d*nsh+sh]; f[ip] += o/dt; }
            }
            lu_solve(&lu, f, x);
            for (int ip = 0; ip < nip; ip++) s->vel[ip][d] = x[ip];
        }
    }
    return 0;
}

/* NavierStokesFV1WithoutStabilization::update, stabilization.cpp:805-850 */
static int stab_none(const ora_params *p, const Geom *g, const double (*stdvel)[3], int bStokes, Stab *s)
{
    s->connected = 0;
    if (!bStokes) if (upwind_compute(p->stab_upwind, g, stdvel, &s->up)) return -1;
    for (int ip = 0; ip < g->nip; ip++) {
        for (int i = 0; i < g->dim; i++) for (int sh = 0; sh < g->nsh; sh++) s->sp[ip][i][sh] = 0;
        for (int sh = 0; sh < g->nsh; sh++) for (int i = 0; i < g->dim; i++) {
            for (int j = 0; j < g->dim; j++) s->sv[ip][i][j][sh] = 0;
            s->sv[ip][i][i][sh] = g->N[ip][sh];
        }
        for (int d = 0; d < 3; d++) s->vel[ip][d] = stdvel[ip][d];
    }
    return 0;
}

static int stab_update(const ora_params *p, const Geom *g, const double *u, const double (*stdvel)[3],
                       int bStokes, const double *uold, double dt, Stab *s)
{
    switch (p->stab) {
    case ORA_STAB_FIELDS: return stab_fields(p, g, u, stdvel, bStokes, uold, dt, s);
    case ORA_STAB_FLOW:   return stab_flow(p, g, u, stdvel, bStokes, uold, dt, s);
    case ORA_STAB_NONE:   return stab_none(p, g, stdvel, bStokes, s);
    }
    return fail("Stabilization has not been set.");     /* fv1/navier_stokes_fv1.cpp:147 */
}

/* prep_elem_loop validation, fv1/navier_stokes_fv1.cpp:136-181 */
static int fv1_validate(const ora_params *p)
{
    if (p->stab < 0 || p->stab > 2) return fail("Stabilization has not been set.");
    if (!p->stokes) {
        if (!p->pac && p->conv_upwind == ORA_UPWIND_NONE) return fail("Upwinding for convective Term in Momentum eq. not set.");
        if (p->stab_upwind == ORA_UPWIND_NONE) return fail("Stabilization has no upwind (UG_NSSTAB_ASSERT: No upwind object).");
    }
    if (!(p->kin_visc == p->kin_visc)) return fail("NavierStokes::prep_elem_loop: Kinematic Viscosity has not been set, but is required.");
    if (!(p->density == p->density)) return fail("NavierStokes::prep_elem_loop: Density has not been set, but is required.");
    return 0;
}

/* peclet_blend, fv1/navier_stokes_fv1.cpp:871-892 */
static double peclet_blend_fv1(double *U, const Geom *g, int ip, const double *stdvel, double visc)
{
    double Pe = vdot(stdvel, g->n[ip], g->dim)/vdot(g->n[ip], g->n[ip], g->dim)
              * vdist(g->x[g->to[ip]], g->x[g->from[ip]], g->dim) / visc;
    double Pe2 = Pe*Pe, w = Pe2/(5.0+Pe2);
    for (int d = 0; d < g->dim; d++) U[d] = w*U[d] + (1.0-w)*stdvel[d];
    return w;
}

#define JL(rf,rsh,cf,csh) Jloc[((rf)*nsh+(rsh))*L + ((cf)*nsh+(csh))]
#define DL(f,sh) dloc[(f)*nsh+(sh)]

/* common prologue of add_jac_A_elem / add_def_A_elem: fv1/navier_stokes_fv1.cpp:261-314, 610-664 */
typedef struct { Geom g; double stdvel[MAXIP][3]; Stab stab; Upw conv; const Upw *upw; int conv_by_stab; } FV1Ctx;

static int fv1_prologue(const ora_params *p, const double *coords, const double *u,
                        const double *sol0, const double *sol1, FV1Ctx *c)
{
    if (fv1_validate(p)) return -1;
    if (geom_update(&c->g, p->elem, coords)) return -1;
    const Geom *g = &c->g; int nsh = g->nsh, dim = g->dim;
    const double *pSol = u, *pOld = NULL; double dt = 0.0;
    if (p->time_dependent) {
        if (!sol0 || !sol1) return fail("NavierStokes::add_jac_A_elem:  Stabilization needs exactly two time points.");
        pSol = sol0; pOld = sol1; dt = p->dt;
    }
    for (int ip = 0; ip < g->nip; ip++) {
        for (int d = 0; d < 3; d++) c->stdvel[ip][d] = 0.0;
        for (int sh = 0; sh < nsh; sh++) for (int d = 0; d < dim; d++) c->stdvel[ip][d] += U_(d,sh)*g->N[ip][sh];
    }
    if (stab_update(p, g, pSol, c->stdvel, p->stokes, pOld, dt, &c->stab)) return -1;
    c->conv_by_stab = 0; c->upw = NULL;
    if (!p->stokes) {
        if (p->pac) c->conv_by_stab = 1;            /* m_spConvStab == m_spStab: not updated twice */
        else {
            /* result identical whether the object is shared with the stab or not */
            if (upwind_compute(p->conv_upwind, g, c->stdvel, &c->conv)) return -1;
            c->upw = &c->conv;
        }
    }
    return 0;
}

/* add_jac_A_elem, fv1/navier_stokes_fv1.cpp:250-595 */
static int fv1_jac_A(const ora_params *p, const FV1Ctx *c, const double *u, double *Jloc)
{
    const Geom *g = &c->g; const Stab *stab = &c->stab, *convStab = &c->stab; const Upw *upwind = c->upw;
    int dim = g->dim, nsh = g->nsh, nip = g->nip, P = dim, L = (dim+1)*nsh;
    for (int ip = 0; ip < nip; ip++) {
        const double visc = VISC(ip), rho = RHOF(ip);
        int f = g->from[ip], t = g->to[ip]; const double *n = g->n[ip];
        for (int sh = 0; sh < nsh; sh++) {
            double flux_sh = -1.0*visc*rho*vdot(g->G[ip][sh], n, dim);
            for (int d1 = 0; d1 < dim; d1++) { JL(d1,f,d1,sh) += flux_sh; JL(d1,t,d1,sh) -= flux_sh; }
            if (!p->laplace)
                for (int d1 = 0; d1 < dim; d1++) for (int d2 = 0; d2 < dim; d2++) {
                    double flux2 = -1.0*visc*rho*g->G[ip][sh][d1]*n[d2];
                    JL(d1,f,d2,sh) += flux2; JL(d1,t,d2,sh) -= flux2;
                }
            for (int d1 = 0; d1 < dim; d1++) { double fl = g->N[ip][sh]*n[d1]; JL(d1,f,P,sh) += fl; JL(d1,t,P,sh) -= fl; }

            if (!p->stokes) {
                double U[3];
                if (upwind) upwind_vel(upwind, g, ip, u, c->stdvel, U);
                else if (c->conv_by_stab) { for (int d = 0; d < 3; d++) U[d] = convStab->vel[ip][d]; }
                else return fail("Cannot find upwind for convective term.");
                double w = 1.0;
                if (p->peclet_blend) w = peclet_blend_fv1(U, g, ip, c->stdvel[ip], visc);
                double prod = vdot(c->stdvel[ip], n, dim)*rho;

                if (c->conv_by_stab) {
                    if (stab->connected) {
                        for (int d1 = 0; d1 < dim; d1++) for (int d2 = 0; d2 < dim; d2++) {
                            double v = prod*w*convStab->sv[ip][d1][d2][sh];
                            JL(d1,f,d2,sh) += v; JL(d1,t,d2,sh) -= v;
                        }
                    } else {
                        for (int d1 = 0; d1 < dim; d1++) { double v = prod*w*convStab->sv[ip][d1][d1][sh]; JL(d1,f,d1,sh) += v; JL(d1,t,d1,sh) -= v; }
                    }
                    for (int d1 = 0; d1 < dim; d1++) { double v = prod*w*convStab->sp[ip][d1][sh]; JL(d1,f,P,sh) += v; JL(d1,t,P,sh) -= v; }
                }
                if (upwind) {
                    double cf = upwind->sh[ip][sh];
                    if (upwind->nonzero_ip) for (int ip2 = 0; ip2 < nip; ip2++) cf += g->N[ip2][sh]*upwind->ip[ip][ip2];
                    cf *= prod*w;
                    for (int d1 = 0; d1 < dim; d1++) { JL(d1,f,d1,sh) += cf; JL(d1,t,d1,sh) -= cf; }
                }
                if (p->peclet_blend) {
                    double v = prod*(1.0-w)*g->N[ip][sh];
                    for (int d1 = 0; d1 < dim; d1++) { JL(d1,f,d1,sh) += v; JL(d1,t,d1,sh) -= v; }
                }
                if (p->exact_jac) {
                    if (c->conv_by_stab) {
                        for (int d1 = 0; d1 < dim; d1++) {
                            for (int d2 = 0; d2 < dim; d2++) {
                                double pv = 0.0;
                                if (stab->connected) for (int k = 0; k < dim; k++) pv += w*convStab->sv[ip][k][d2][sh]*n[k];
                                else pv = convStab->sv[ip][d1][d1][sh]*n[d1];          /* quirk :494-496 */
                                pv *= p->exact_jac*rho;
                                JL(d1,f,d2,sh) += pv*U[d1]; JL(d1,t,d2,sh) -= pv*U[d1];
                            }
                            double pp = 0.0;
                            for (int k = 0; k < dim; k++) pp += convStab->sp[ip][k][sh]*n[k];
                            pp *= p->exact_jac*rho;
                            JL(d1,f,P,sh) += pp*U[d1]; JL(d1,t,P,sh) -= pp*U[d1];
                        }
                    }
                    if (upwind) {
                        for (int d1 = 0; d1 < dim; d1++) for (int d2 = 0; d2 < dim; d2++) {
                            double pv = w*upwind->sh[ip][sh]*n[d2]*rho;                 /* quirk :528-529 */
                            JL(d1,f,d2,sh) += pv*U[d1]; JL(d1,t,d2,sh) -= pv*U[d1];
                        }
                    }
                    if (p->peclet_blend) {
                        for (int d1 = 0; d1 < dim; d1++) for (int d2 = 0; d2 < dim; d2++) {
                            double v = U[d1]*(1.0-w)*g->N[ip][sh]*n[d2]*rho;          /* quirk :542-545 */
                            JL(d1,f,d2,sh) += v; JL(d1,t,d2,sh) -= v;
                        }
                    }
                }
            }
            /* continuity */
            if (stab->connected) {
                for (int d1 = 0; d1 < dim; d1++) {
                    double cv = 0.0;
                    for (int d2 = 0; d2 < dim; d2++) cv += stab->sv[ip][d2][d1][sh]*n[d2]*rho;
                    JL(P,f,d1,sh) += cv; JL(P,t,d1,sh) -= cv;
                }
            } else {
                for (int d1 = 0; d1 < dim; d1++) { double cv = stab->sv[ip][d1][d1][sh]*n[d1]*rho; JL(P,f,d1,sh) += cv; JL(P,t,d1,sh) -= cv; }
            }
            double cp = 0.0;
            for (int d1 = 0; d1 < dim; d1++) cp += stab->sp[ip][d1][sh]*n[d1]*rho;
            JL(P,f,P,sh) += cp; JL(P,t,P,sh) -= cp;
        }
    }
    return 0;
}

/* add_def_A_elem, fv1/navier_stokes_fv1.cpp:597-778 */
static int fv1_def_A(const ora_params *p, const FV1Ctx *c, const double *u, double *dloc)
{
    const Geom *g = &c->g; const Stab *stab = &c->stab; const Upw *upwind = c->upw;
    int dim = g->dim, nsh = g->nsh, nip = g->nip, P = dim;
    for (int ip = 0; ip < nip; ip++) {
        const double visc = VISC(ip), rho = RHOF(ip);
        int f = g->from[ip], t = g->to[ip]; const double *n = g->n[ip];
        double gradVel[3][3], diff[3];
        for (int d1 = 0; d1 < dim; d1++) for (int d2 = 0; d2 < dim; d2++) {
            gradVel[d1][d2] = 0.0;
            for (int sh = 0; sh < nsh; sh++) gradVel[d1][d2] += g->G[ip][sh][d2]*U_(d1,sh);
        }
        for (int d1 = 0; d1 < dim; d1++) { diff[d1] = 0; for (int d2 = 0; d2 < dim; d2++) diff[d1] += gradVel[d1][d2]*n[d2]; }
        if (!p->laplace) for (int d1 = 0; d1 < dim; d1++) for (int d2 = 0; d2 < dim; d2++) diff[d1] += gradVel[d2][d1]*n[d2];
        for (int d1 = 0; d1 < dim; d1++) diff[d1] *= (-1.0)*visc*rho;
        for (int d1 = 0; d1 < dim; d1++) { DL(d1,f) += diff[d1]; DL(d1,t) -= diff[d1]; }
        if (!p->stokes) {
            double U[3];
            if (upwind) upwind_vel(upwind, g, ip, u, c->stdvel, U);
            else if (c->conv_by_stab) { for (int d = 0; d < 3; d++) U[d] = stab->vel[ip][d]; }
            else return fail("Cannot find upwind for convective term.");
            if (p->peclet_blend) peclet_blend_fv1(U, g, ip, c->stdvel[ip], visc);
            double prod = vdot(c->stdvel[ip], n, dim)*rho;
            for (int d1 = 0; d1 < dim; d1++) { DL(d1,f) += U[d1]*prod; DL(d1,t) -= U[d1]*prod; }
        }
        double pr = 0.0;
        for (int sh = 0; sh < nsh; sh++) pr += g->N[ip][sh]*U_(P,sh);
        for (int d1 = 0; d1 < dim; d1++) { DL(d1,f) += pr*n[d1]; DL(d1,t) -= pr*n[d1]; }
        double cont = vdot(stab->vel[ip], n, dim)*rho;
        DL(P,f) += cont; DL(P,t) -= cont;
    }
    return 0;
}

int ora_fv1_elem(const ora_params *p, const double *coords, const double *u,
                 const double *sol0, const double *sol1, int what, double *Jloc, double *dloc)
{
    int nsh = ora_elem_nsh(p->elem), dim = ora_elem_dim(p->elem);
    if (nsh < 0) return fail("unknown element type");
    int L = (dim+1)*nsh;
    FV1Ctx c;
    if (what & (ORA_JAC_A|ORA_DEF_A)) {
        if (fv1_prologue(p, coords, u, sol0, sol1, &c)) return -1;
        if ((what & ORA_JAC_A) && fv1_jac_A(p, &c, u, Jloc)) return -1;
        if ((what & ORA_DEF_A) && fv1_def_A(p, &c, u, dloc)) return -1;
    } else if (geom_update(&c.g, p->elem, coords)) return -1;
    /* add_jac_M_elem :781-808, add_def_M_elem :811-838, add_rhs_elem :841-869 */
    if (what & ORA_JAC_M) for (int sh = 0; sh < nsh; sh++) for (int d1 = 0; d1 < dim; d1++) JL(d1,sh,d1,sh) += c.g.vol[sh]*RHOV(sh);
    if (what & ORA_DEF_M) for (int sh = 0; sh < nsh; sh++) for (int d1 = 0; d1 < dim; d1++) DL(d1,sh) += U_(d1,sh)*c.g.vol[sh]*RHOV(sh);
    if ((what & ORA_RHS) && HAS_SRCV)
        for (int sh = 0; sh < nsh; sh++) for (int d1 = 0; d1 < dim; d1++) DL(d1,sh) += SRCV(sh,d1)*c.g.vol[sh]*RHOV(sh);
    return 0;
}

int ora_fv1_stab(const ora_params *p, const double *coords, const double *u, const double *sol0,
                 const double *sol1, double *stab_vel, double *shape_vel, double *shape_p)
{
    FV1Ctx c; if (fv1_prologue(p, coords, u, sol0, sol1, &c)) return -1;
    int dim = c.g.dim, nsh = c.g.nsh, nip = c.g.nip;
    for (int ip = 0; ip < nip; ip++) for (int d = 0; d < dim; d++) {
        stab_vel[ip*dim+d] = c.stab.vel[ip][d];
        for (int k = 0; k < nsh; k++) {
            shape_p[(ip*dim+d)*nsh+k] = c.stab.sp[ip][d][k];
            for (int d2 = 0; d2 < dim; d2++) shape_vel[((ip*dim+d)*dim+d2)*nsh+k] = c.stab.sv[ip][d][d2][k];
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * CRFVGeometry restatement (ugcore fvcr_geom.cpp -- App. B-3, our spec). Simplices only
 * (regular grids; hanging-node HCR branches are out of scope).
 * SCV per side (node_id = side), SCVF per (dim-2)-object: 2-D per corner, 3-D per edge,
 * spanned by that object and the barycentre; from/to = the two sides sharing the object
 * (lower side index = from); normal oriented from -> to.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int elem, dim, nsh, nip, nco;
    double x[MAXSH][3], bary[3];
    int from[MAXIP], to[MAXIP];
    double n[MAXIP][3], xip[MAXIP][3], lip[MAXIP][3], N[MAXIP][6], G[MAXIP][6][3];
    double scv_n[6][3], scv_xip[6][3], vol[6];
} CRGeom;

/* Crouzeix-Raviart shapes on simplices: 1 - dim*lambda_opposite(side) */
static int cr_opposite(int elem, int side) {
    if (elem == ORA_TRI) return (side+2)%3;
    static const int opp[4] = {3,0,1,2};
    return opp[side];
}
/* Crouzeix-Raviart (rotated bi-/trilinear, Rannacher-Turek point-value variant) shapes on quadrilaterals / hexahedra:
   span {1, x, y, x^2 - y^2} resp. {1, x, y, z, x^2 - y^2, y^2 - z^2}, nodal at the side centres (ugcore
   CrouzeixRaviartLSFS<ReferenceQuadrilateral / ReferenceHexahedron> -- our spec). The coefficients are obtained once by
   inverting the 4 x 4 / 6 x 6 generalised Vandermonde matrix. */
static double g_crq_coef[2][6][6]; static int g_crq_ready[2];
static void crq_basis(int dim, const double *x, double *b, double (*db)[3])
{
    if (dim == 2) {
        b[0] = 1; b[1] = x[0]; b[2] = x[1]; b[3] = x[0]*x[0]-x[1]*x[1];
        if (db) { double t[4][3] = {{0,0,0},{1,0,0},{0,1,0},{2*x[0],-2*x[1],0}}; memcpy(db, t, sizeof t); }
    } else {
        b[0] = 1; b[1] = x[0]; b[2] = x[1]; b[3] = x[2]; b[4] = x[0]*x[0]-x[1]*x[1]; b[5] = x[1]*x[1]-x[2]*x[2];
        if (db) { double t[6][3] = {{0,0,0},{1,0,0},{0,1,0},{0,0,1},{2*x[0],-2*x[1],0},{0,2*x[1],-2*x[2]}}; memcpy(db, t, sizeof t); }
    }
}
static const RefElem *get_ref(int elem);
static void crq_init(int elem)
{
    const int w = elem == ORA_QUAD ? 0 : 1;
    if (g_crq_ready[w]) return;
#pragma omp critical(ora_crq_init)
    if (!g_crq_ready[w]) {
        const RefElem *r = get_ref(elem); const int n = r->nside, dim = r->dim;
        double V[6][12];                                         /* [V | I] -> [I | V^-1], Gauss-Jordan with partial pivoting */
        for (int sd = 0; sd < n; sd++) {
            double c[3]; avg_pts(c, r->corner, r->side[sd], r->side_n[sd], dim); if (dim == 2) c[2] = 0;
            double b[6]; crq_basis(dim, c, b, NULL);
            for (int j = 0; j < n; j++) { V[sd][j] = b[j]; V[sd][n+j] = sd == j; }
        }
        for (int k = 0; k < n; k++) {
            int pv = k; for (int i = k+1; i < n; i++) if (fabs(V[i][k]) > fabs(V[pv][k])) pv = i;
            for (int j = 0; j < 2*n; j++) { double t = V[k][j]; V[k][j] = V[pv][j]; V[pv][j] = t; }
            double piv = V[k][k]; for (int j = 0; j < 2*n; j++) V[k][j] /= piv;
            for (int i = 0; i < n; i++) if (i != k) { double f = V[i][k]; for (int j = 0; j < 2*n; j++) V[i][j] -= f*V[k][j]; }
        }
        /* N_s(x) = sum_j coef[s][j] b_j(x) with V coef^T = I  ->  coef[s][j] = (V^-1)[j][s] */
        for (int sd = 0; sd < n; sd++) for (int j = 0; j < n; j++) g_crq_coef[w][sd][j] = V[j][n+sd];
        g_crq_ready[w] = 1;
    }
}
static void cr_shapes(int elem, const double *xi, double *N, double (*dN)[3])
{
    if (elem == ORA_QUAD || elem == ORA_HEX) {
        crq_init(elem);
        const int w = elem == ORA_QUAD ? 0 : 1, n = elem == ORA_QUAD ? 4 : 6, dim = elem == ORA_QUAD ? 2 : 3;
        double b[6], db[6][3];
        crq_basis(dim, xi, b, db);
        for (int sd = 0; sd < n; sd++) {
            double v = 0, g[3] = {0,0,0};
            for (int j = 0; j < n; j++) { v += g_crq_coef[w][sd][j]*b[j]; for (int d = 0; d < 3; d++) g[d] += g_crq_coef[w][sd][j]*db[j][d]; }
            N[sd] = v; if (dN) for (int d = 0; d < 3; d++) dN[sd][d] = g[d];
        }
        return;
    }
    double lam[4], dl[4][3]; int nco = elem == ORA_TRI ? 3 : 4, dim = elem == ORA_TRI ? 2 : 3;
    lagrange_shapes(elem, xi, lam, dl);
    for (int s = 0; s < nco; s++) {
        int o = cr_opposite(elem, s);
        N[s] = 1.0 - dim*lam[o];
        if (dN) for (int d = 0; d < 3; d++) dN[s][d] = -dim*dl[o][d];
    }
}

static int cr_geom_update(CRGeom *g, int elem, const double *coords)
{
    if (elem != ORA_TRI && elem != ORA_TET && elem != ORA_QUAD && elem != ORA_HEX) return fail("CRFVGeometry oracle: tri / quad / tet / hex only");
    const int simplex = elem == ORA_TRI || elem == ORA_TET;
    const RefElem *r = get_ref(elem);
    int dim = r->dim, nco = r->nsh, nside = r->nside;
    g->elem = elem; g->dim = dim; g->nco = nco; g->nsh = nside; g->nip = dim == 2 ? nco : r->nedge;
    for (int i = 0; i < nco; i++) { for (int d = 0; d < dim; d++) g->x[i][d] = coords[i*dim+d]; for (int d = dim; d < 3; d++) g->x[i][d] = 0; }
    int all[MAXSH]; for (int i = 0; i < nco; i++) all[i] = i;
    avg_pts(g->bary, g->x, all, nco, dim);
    double lbary[3]; avg_pts(lbary, r->corner, all, nco, dim);
    /* JTInv (constant on simplices) */
    double JT[3][3] = {{0}}, JTinv[3][3] = {{0}}, lam[MAXSH], dl[MAXSH][3], z[3] = {0,0,0};
    lagrange_shapes(elem, z, lam, dl);
    for (int i = 0; i < dim; i++) for (int j = 0; j < dim; j++) { double s = 0; for (int k = 0; k < nco; k++) s += dl[k][i]*g->x[k][j]; JT[i][j] = s; }
    double det = mat_inverse(dim, JT, JTinv);
    if (!(fabs(det) > 0)) return fail("CRFVGeometry: singular element Jacobian");
    double elemvol = fabs(det) / (dim == 2 ? 2.0 : 6.0);
    /* SCVs */
    for (int s = 0; s < nside; s++) {
        avg_pts(g->scv_xip[s], g->x, r->side[s], r->side_n[s], dim);
        g->vol[s] = elemvol / nside;                       /* cone side<->barycentre of a simplex */
        double nn[3] = {0,0,0};
        if (dim == 2) {
            const double *a = g->x[r->side[s][0]], *b = g->x[r->side[s][1]];
            nn[0] = b[1]-a[1]; nn[1] = -(b[0]-a[0]);
            if (!simplex)                                  /* triangle (side, barycentre) */
                g->vol[s] = 0.5*fabs((b[0]-a[0])*(g->bary[1]-a[1]) - (b[1]-a[1])*(g->bary[0]-a[0]));
        } else if (r->side_n[s] == 4) {
            /* quadrilateral side: area vector 0.5 (c2-c0) x (c3-c1); SCV = pyramid (side, barycentre), its volume as the two
               tetrahedra of the side split along its diagonal 0-2 (the split of the ray search) */
            const double *c0 = g->x[r->side[s][0]], *c1 = g->x[r->side[s][1]], *c2 = g->x[r->side[s][2]], *c3 = g->x[r->side[s][3]];
            double d1[3], d2[3], a1[3], a2[3], a3[3], t[3];
            for (int d = 0; d < 3; d++) { d1[d] = c2[d]-c0[d]; d2[d] = c3[d]-c1[d]; }
            vcross(nn, d1, d2); for (int d = 0; d < 3; d++) nn[d] *= 0.5;
            for (int d = 0; d < 3; d++) { a1[d] = c1[d]-c0[d]; a2[d] = c2[d]-c0[d]; a3[d] = g->bary[d]-c0[d]; }
            vcross(t, a1, a2); double v = fabs(vdot(t, a3, 3));
            for (int d = 0; d < 3; d++) a1[d] = c3[d]-c0[d];
            vcross(t, a2, a1); v += fabs(vdot(t, a3, 3));
            g->vol[s] = v/6.0;
        } else {
            double e1[3], e2[3];
            for (int d = 0; d < 3; d++) { e1[d] = g->x[r->side[s][1]][d]-g->x[r->side[s][0]][d]; e2[d] = g->x[r->side[s][2]][d]-g->x[r->side[s][0]][d]; }
            vcross(nn, e1, e2); for (int d = 0; d < 3; d++) nn[d] *= 0.5;
        }
        double out[3]; for (int d = 0; d < 3; d++) out[d] = g->scv_xip[s][d] - g->bary[d];
        double sg = vdot(nn, out, dim) < 0 ? -1.0 : 1.0;             /* outward */
        for (int d = 0; d < 3; d++) g->scv_n[s][d] = sg*nn[d];
    }
    /* SCVFs */
    for (int ip = 0; ip < g->nip; ip++) {
        int obj[2], nobj = dim == 2 ? 1 : 2;
        if (dim == 2) obj[0] = ip; else { obj[0] = r->edge[ip][0]; obj[1] = r->edge[ip][1]; }
        int sides[2], ns = 0;
        for (int s = 0; s < nside && ns < 2; s++) {
            int hit = 0;
            for (int k = 0; k < r->side_n[s]; k++) for (int q = 0; q < nobj; q++) if (r->side[s][k] == obj[q]) hit++;
            if (hit == nobj) sides[ns++] = s;
        }
        g->from[ip] = sides[0]; g->to[ip] = sides[1];
        double nn[3] = {0,0,0};
        for (int d = 0; d < 3; d++) { g->xip[ip][d] = 0; g->lip[ip][d] = 0; }
        for (int q = 0; q < nobj; q++) for (int d = 0; d < dim; d++) { g->xip[ip][d] += g->x[obj[q]][d]; g->lip[ip][d] += r->corner[obj[q]][d]; }
        for (int d = 0; d < dim; d++) { g->xip[ip][d] = (g->xip[ip][d] + g->bary[d])/(nobj+1); g->lip[ip][d] = (g->lip[ip][d] + lbary[d])/(nobj+1); }
        if (dim == 2) { const double *a = g->x[obj[0]]; nn[0] = g->bary[1]-a[1]; nn[1] = -(g->bary[0]-a[0]); }
        else {
            double e1[3], e2[3];
            for (int d = 0; d < 3; d++) { e1[d] = g->x[obj[1]][d]-g->x[obj[0]][d]; e2[d] = g->bary[d]-g->x[obj[0]][d]; }
            vcross(nn, e1, e2); for (int d = 0; d < 3; d++) nn[d] *= 0.5;
        }
        double ft[3]; for (int d = 0; d < 3; d++) ft[d] = g->scv_xip[g->to[ip]][d] - g->scv_xip[g->from[ip]][d];
        double sg = vdot(nn, ft, dim) < 0 ? -1.0 : 1.0;
        for (int d = 0; d < 3; d++) g->n[ip][d] = sg*nn[d];
        double dN[6][3];
        cr_shapes(elem, g->lip[ip], g->N[ip], dN);
        if (!simplex) {                                    /* JTInv at the local ip (bi-/trilinear element map) */
            lagrange_shapes(elem, g->lip[ip], lam, dl);
            for (int i = 0; i < dim; i++) for (int j = 0; j < dim; j++) { double sm = 0; for (int k = 0; k < nco; k++) sm += dl[k][i]*g->x[k][j]; JT[i][j] = sm; }
            if (!(fabs(mat_inverse(dim, JT, JTinv)) > 0)) return fail("CRFVGeometry: singular element Jacobian");
        }
        for (int k = 0; k < nside; k++) {
            for (int j = 0; j < 3; j++) g->G[ip][k][j] = 0;
            for (int j = 0; j < dim; j++) { double s = 0; for (int i = 0; i < dim; i++) s += JTinv[j][i]*dN[k][i]; g->G[ip][k][j] = s; }
        }
    }
    return 0;
}

int ora_cr_geometry(int elem, const double *coords, ora_cr_geom *out)
{
    CRGeom g; if (cr_geom_update(&g, elem, coords)) return -1;
    memset(out, 0, sizeof *out);
    out->dim = g.dim; out->nsh = g.nsh; out->nip = g.nip; out->nco = g.nco;
    for (int ip = 0; ip < g.nip; ip++) {
        out->from[ip] = g.from[ip]; out->to[ip] = g.to[ip];
        for (int d = 0; d < 3; d++) { out->normal[ip][d] = g.n[ip][d]; out->xip[ip][d] = g.xip[ip][d]; out->lip[ip][d] = g.lip[ip][d]; }
        for (int k = 0; k < g.nsh; k++) { out->shape[ip][k] = g.N[ip][k]; for (int d = 0; d < 3; d++) out->ggrad[ip][k][d] = g.G[ip][k][d]; }
    }
    for (int s = 0; s < g.nsh; s++) { out->vol[s] = g.vol[s]; for (int d = 0; d < 3; d++) { out->scv_normal[s][d] = g.scv_n[s][d]; out->scv_xip[s][d] = g.scv_xip[s][d]; } }
    return 0;
}

/* CR-geometry upwinds: upwind.cpp:82-104 (No), :174-213 (Full), :432-499 (Skewed), :577-636 (LPS) */
static int cr_upwind(int type, const CRGeom *g, const double (*vel)[3], Upw *u)
{
    const RefElem *r = get_ref(g->elem);
    int nip = g->nip, nsh = g->nsh, dim = g->dim;
    u->nonzero_ip = 0;
    for (int i = 0; i < MAXIP; i++) for (int j = 0; j < MAXIP; j++) u->ip[i][j] = NAN;
    for (int ip = 0; ip < nip; ip++) {
        u->len[ip] = NAN;                                 /* FVCR never reads the conv length */
        if (type == ORA_UPWIND_NO) { for (int sh = 0; sh < nsh; sh++) u->sh[ip][sh] = g->N[ip][sh]; continue; }
        for (int sh = 0; sh < nsh; sh++) u->sh[ip][sh] = 0.0;
        if (type == ORA_UPWIND_FULL) {
            double flux = vdot(g->n[ip], vel[ip], dim);
            int s = flux > 0.0 ? g->from[ip] : g->to[ip];
            u->sh[ip][s] = 1.0; u->len[ip] = vdist(g->xip[ip], g->scv_xip[s], dim);
            continue;
        }
        if (type == ORA_UPWIND_SKEWED) { if (sqrt(vdot(vel[ip], vel[ip], dim)) < 1e-14) continue; }
        else if (type == ORA_UPWIND_LPS) { if (sqrt(vdot(vel[ip], vel[ip], dim)) == 0.0) continue; }
        else return fail("No update function registered for Geometry (upwind has no CR overload)"); /* upwind_interface.h:316-318 */
        int side; double gc[3], lc[3], N[6];
        if (!side_ray_intersection(r, g->x, g->xip[ip], vel[ip], 0, &side, gc, lc))
            return fail("GetSkewedUpwindShapes: Cannot find cut side.");
        cr_shapes(g->elem, lc, N, NULL);
        if (type == ORA_UPWIND_SKEWED) {
            double max = -1000; int maxind = 0;
            for (int sh = 0; sh < nsh; sh++) if (N[sh] > max) { max = N[sh]; maxind = sh; }
            u->sh[ip][maxind] = 1; u->len[ip] = vdist(g->xip[ip], g->scv_xip[maxind], dim);
        } else {
            for (int sh = 0; sh < nsh; sh++) u->sh[ip][sh] = N[sh];
            u->len[ip] = vdist(g->xip[ip], gc, dim);
        }
    }
    return 0;
}

/* FVCR element routines, fvcr/navier_stokes_fvcr.cpp:244-759.
   Local dofs: velocity (d, side) at d*nsh+side, pressure at dim*nsh. L = dim*nsh+1 */
int ora_fvcr_elem(const ora_params *p, const double *coords, const double *u, int what, double *Jloc, double *dloc)
{
    CRGeom G; if (cr_geom_update(&G, p->elem, coords)) return -1;
    const CRGeom *g = &G;
    int dim = g->dim, nsh = g->nsh, nip = g->nip, L = dim*nsh+1, PI = dim*nsh;
    double visc = p->kin_visc, rho = p->density;
#define JC(r,c) Jloc[(r)*L+(c)]
    if (!p->stokes && p->conv_upwind == ORA_UPWIND_NONE && (what & (ORA_JAC_A|ORA_DEF_A)))
        return fail("Upwinding for convective Term in Momentum eq. not set.");     /* :158-159 */
    double stdvel[MAXIP][3]; Upw up; Geom gv;            /* gv only carries sizes for upwind_vel */
    gv.nsh = nsh; gv.nip = nip; gv.dim = dim;
    if (what & (ORA_JAC_A|ORA_DEF_A)) {
        for (int ip = 0; ip < nip; ip++) {
            for (int d = 0; d < 3; d++) stdvel[ip][d] = 0;
            for (int sh = 0; sh < nsh; sh++) for (int d = 0; d < dim; d++) stdvel[ip][d] += u[d*nsh+sh]*g->N[ip][sh];
        }
    }
    if (what & ORA_JAC_A) {
        if (!p->stokes) if (cr_upwind(p->conv_upwind, g, stdvel, &up)) return -1;
        for (int ip = 0; ip < nip; ip++) {
            int f = g->from[ip], t = g->to[ip]; const double *n = g->n[ip];
            for (int sh = 0; sh < nsh; sh++) {
                double flux_sh = -1.0*visc*rho*vdot(g->G[ip][sh], n, dim);
                for (int d1 = 0; d1 < dim; d1++) { JC(d1*nsh+f, d1*nsh+sh) += flux_sh; JC(d1*nsh+t, d1*nsh+sh) -= flux_sh; }
                if (!p->laplace) for (int d1 = 0; d1 < dim; d1++) for (int d2 = 0; d2 < dim; d2++) {
                    double fl = -1.0*visc*rho*g->G[ip][sh][d1]*n[d2];
                    JC(d1*nsh+f, d2*nsh+sh) += fl; JC(d1*nsh+t, d2*nsh+sh) -= fl;
                }
                if (p->grad_div > 0) for (int d1 = 0; d1 < dim; d1++) for (int d2 = 0; d2 < dim; d2++) {
                    double sf = p->grad_div*g->G[ip][sh][d2]*n[d1];
                    JC(d1*nsh+f, d2*nsh+sh) -= sf; JC(d1*nsh+t, d2*nsh+sh) += sf;
                }
                if (!p->stokes) {
                    double U[3]; upwind_vel(&up, &gv, ip, u, stdvel, U);
                    double w = 1.0;
                    if (p->peclet_blend) {                           /* :244-265 */
                        double Pe = vdot(stdvel[ip], n, dim)/vdot(n, n, dim)*vdist(g->scv_xip[t], g->scv_xip[f], dim)/visc;
                        double Pe2 = Pe*Pe; w = Pe2/(5.0+Pe2);
                        for (int d = 0; d < dim; d++) U[d] = w*U[d] + (1.0-w)*stdvel[ip][d];
                    }
                    double prod = vdot(stdvel[ip], n, dim)*rho;
                    double cf = up.sh[ip][sh];
                    if (up.nonzero_ip) for (int ip2 = 0; ip2 < nip; ip2++) cf += g->N[ip2][sh]*up.ip[ip][ip2];
                    cf *= prod*w;
                    for (int d1 = 0; d1 < dim; d1++) { JC(d1*nsh+f, d1*nsh+sh) += cf; JC(d1*nsh+t, d1*nsh+sh) -= cf; }
                    if (p->peclet_blend) {
                        double v = prod*(1.0-w)*g->N[ip][sh];
                        for (int d1 = 0; d1 < dim; d1++) { JC(d1*nsh+f, d1*nsh+sh) += v; JC(d1*nsh+t, d1*nsh+sh) -= v; }
                    }
                    if (p->exact_jac) {
                        for (int d1 = 0; d1 < dim; d1++) for (int d2 = 0; d2 < dim; d2++) {
                            double pv = p->exact_jac*rho*stdvel[ip][d1]*n[d2]*g->N[ip][sh];   /* :425-426 */
                            JC(d1*nsh+f, d2*nsh+sh) += pv; JC(d1*nsh+t, d2*nsh+sh) -= pv;
                        }
                        if (p->peclet_blend) for (int d1 = 0; d1 < dim; d1++) for (int d2 = 0; d2 < dim; d2++) {
                            double v = U[d1]*(1.0-w)*g->N[ip][sh]*n[d2]*rho*p->exact_jac;
                            JC(d1*nsh+f, d2*nsh+sh) += v; JC(d1*nsh+t, d2*nsh+sh) -= v;
                        }
                    }
                }
            }
            for (int d1 = 0; d1 < dim; d1++) { JC(d1*nsh+f, PI) += n[d1]; JC(d1*nsh+t, PI) -= n[d1]; }
        }
        for (int sh = 0; sh < nsh; sh++) for (int d1 = 0; d1 < dim; d1++) JC(PI, d1*nsh+sh) += g->scv_n[sh][d1];
    }
    if (what & ORA_DEF_A) {
        if (!p->stokes && p->defect_upwind) if (cr_upwind(p->conv_upwind, g, stdvel, &up)) return -1;
        for (int ip = 0; ip < nip; ip++) {
            int f = g->from[ip], t = g->to[ip]; const double *n = g->n[ip];
            double gradVel[3][3], diff[3];
            for (int d1 = 0; d1 < dim; d1++) for (int d2 = 0; d2 < dim; d2++) {
                gradVel[d1][d2] = 0.0;
                for (int sh = 0; sh < nsh; sh++) gradVel[d1][d2] += g->G[ip][sh][d2]*u[d1*nsh+sh];
            }
            for (int d1 = 0; d1 < dim; d1++) { diff[d1] = 0; for (int d2 = 0; d2 < dim; d2++) diff[d1] += gradVel[d1][d2]*n[d2]; }
            if (!p->laplace) for (int d1 = 0; d1 < dim; d1++) for (int d2 = 0; d2 < dim; d2++) diff[d1] += gradVel[d2][d1]*n[d2];
            for (int d1 = 0; d1 < dim; d1++) { diff[d1] *= (-1.0)*visc*rho; dloc[d1*nsh+f] += diff[d1]; dloc[d1*nsh+t] -= diff[d1]; }
            if (p->grad_div > 0) for (int sh = 0; sh < nsh; sh++) for (int d1 = 0; d1 < dim; d1++) for (int d2 = 0; d2 < dim; d2++) {
                double sf = p->grad_div*g->G[ip][sh][d2]*u[d2*nsh+sh]*n[d1];
                dloc[d1*nsh+f] -= sf; dloc[d1*nsh+t] += sf;
            }
            if (!p->stokes) {
                double prod = vdot(stdvel[ip], n, dim)*rho;
                if (p->defect_upwind) {
                    double U[3]; upwind_vel(&up, &gv, ip, u, stdvel, U);
                    if (p->peclet_blend) {
                        double Pe = vdot(stdvel[ip], n, dim)/vdot(n, n, dim)*vdist(g->scv_xip[t], g->scv_xip[f], dim)/visc;
                        double Pe2 = Pe*Pe, w = Pe2/(5.0+Pe2);
                        for (int d = 0; d < dim; d++) U[d] = w*U[d] + (1.0-w)*stdvel[ip][d];
                    }
                    for (int d1 = 0; d1 < dim; d1++) { dloc[d1*nsh+f] += U[d1]*prod; dloc[d1*nsh+t] -= U[d1]*prod; }
                } else
                    for (int d1 = 0; d1 < dim; d1++) { dloc[d1*nsh+f] += stdvel[ip][d1]*prod; dloc[d1*nsh+t] -= stdvel[ip][d1]*prod; }
            }
            double pr = u[PI];
            for (int d1 = 0; d1 < dim; d1++) { dloc[d1*nsh+f] += pr*n[d1]; dloc[d1*nsh+t] -= pr*n[d1]; }
        }
        for (int sh = 0; sh < nsh; sh++) for (int d1 = 0; d1 < dim; d1++) dloc[PI] += g->scv_n[sh][d1]*u[d1*nsh+sh];
    }
    if (what & ORA_JAC_M) for (int sh = 0; sh < nsh; sh++) for (int d1 = 0; d1 < dim; d1++) JC(d1*nsh+sh, d1*nsh+sh) += g->vol[sh]*rho;
    if (what & ORA_DEF_M) for (int sh = 0; sh < nsh; sh++) for (int d1 = 0; d1 < dim; d1++) dloc[d1*nsh+sh] += u[d1*nsh+sh]*g->vol[sh]*rho;
    if ((what & ORA_RHS) && p->has_source)                 /* no density factor, :757 */
        for (int sh = 0; sh < nsh; sh++) for (int d1 = 0; d1 < dim; d1++) dloc[d1*nsh+sh] += p->source[d1]*g->vol[sh];
#undef JC
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * turbulent viscosity (Smagorinsky) as per-ip kinematic viscosity, and diagnostics
 * ---------------------------------------------------------------------------------------- */
/* FV1SmagorinskyTurbViscData (fv1/turbulent_viscosity_fv1.h:200-383):
 *  - assembleDeformationTensor (fv1/turbulent_viscosity_fv1_impl.h:504-616): per node a
 *      D_a = 1/vol_a [ sum_scvf +-(1/2)(u_ip n^T + n u_ip^T)  +  sum_{BF in the turbulence-zero subsets} (1/2)(u_a n^T + n u_a^T) ],
 *      u_ip = sum_sh shape(ip, sh) u_sh, + at scvf.from(), - at scvf.to(); vol_a = sum of the SCV volumes;
 *  - update (:819-852): nu_t(a) = c delta^2 FNorm(D_a), delta = vol_a^(1/dim), FNorm = sqrt(2 sum D_ij^2) (:755-762); nodes of the
 *    turbulence-zero subsets keep 0;
 *  - evaluate (turbulent_viscosity_fv1.h:321-379): value(ip) = sum_sh N_sh(ip) nu_t(vertex sh) + kinematic viscosity.
 * u [n_node][dim+1]; zero_node [n_node] flags (may be NULL); (belem, bside) the turbulence-zero boundary sides.
 * out: nu_t [n_node], ip_visc [n_elem][nip] (either may be NULL). */
int ora_fv1_smagorinsky(int elem, int64_t n_elem, int64_t n_node, const int32_t *conn, const double *coords, const double *u,
                        double c, double kin_visc, int64_t n_bside, const int32_t *belem, const int32_t *bside,
                        const uint8_t *zero_node, double *nu_t, double *ip_visc)
{
    const RefElem *r = get_ref(elem);
    if (!r) return fail("ora_fv1_smagorinsky: unknown element type");
    const int dim = r->dim, nsh = r->nsh, nip = r->nedge, nf = dim + 1;
    double *D = calloc((size_t)n_node * 9, sizeof *D), *vol = calloc((size_t)n_node, sizeof *vol), *nt = calloc((size_t)n_node, sizeof *nt);
    int rc = 0;
    for (int64_t e = 0; e < n_elem && !rc; e++) {
        double xc[MAXSH*3]; Geom g;
        for (int k = 0; k < nsh; k++) for (int d = 0; d < dim; d++) xc[k*dim+d] = coords[(int64_t)conn[e*nsh+k]*dim+d];
        if ((rc = geom_update(&g, elem, xc))) break;
        for (int k = 0; k < nsh; k++) vol[conn[e*nsh+k]] += g.vol[k];
        for (int ip = 0; ip < nip; ip++) {
            double v[3] = {0,0,0};
            for (int k = 0; k < nsh; k++) for (int d = 0; d < dim; d++) v[d] += g.N[ip][k] * u[(int64_t)conn[e*nsh+k]*nf+d];
            const int64_t a = conn[e*nsh+g.from[ip]], b = conn[e*nsh+g.to[ip]];
            for (int i = 0; i < dim; i++) for (int j = 0; j < dim; j++) {
                const double f = 0.5 * (v[i] * g.n[ip][j] + v[j] * g.n[ip][i]);
                D[a*9+i*3+j] += f; D[b*9+i*3+j] -= f;
            }
        }
    }
    for (int64_t b = 0; b < n_bside && !rc; b++) {
        const int64_t e = belem[b];
        double xc[MAXSH*3];
        for (int k = 0; k < nsh; k++) for (int d = 0; d < dim; d++) xc[k*dim+d] = coords[(int64_t)conn[e*nsh+k]*dim+d];
        for (int j = 0; j < r->side_n[bside[b]] && !rc; j++) {
            BFace bf;
            if ((rc = bf_update(&bf, elem, xc, bside[b], j))) break;
            const int64_t a = conn[e*nsh+bf.node_id];
            for (int i = 0; i < dim; i++) for (int jj = 0; jj < dim; jj++)
                D[a*9+i*3+jj] += 0.5 * (u[a*nf+i] * bf.n[jj] + u[a*nf+jj] * bf.n[i]);
        }
    }
    for (int64_t a = 0; a < n_node && !rc; a++) {
        if (!(vol[a] > 0) || (zero_node && zero_node[a])) { nt[a] = 0.0; continue; }
        double s = 0;
        for (int i = 0; i < dim; i++) for (int j = 0; j < dim; j++) { const double t = D[a*9+i*3+j] / vol[a]; s += t * t; }
        const double delta = pow(vol[a], 1.0 / dim);
        nt[a] = c * delta * delta * sqrt(2.0 * s);
    }
    if (!rc && nu_t) memcpy(nu_t, nt, sizeof(double) * (size_t)n_node);
    if (!rc && ip_visc)
        for (int64_t e = 0; e < n_elem; e++) for (int ip = 0; ip < nip; ip++) {
            double s = 0; for (int k = 0; k < nsh; k++) s += r->shape_ip[ip][k] * nt[conn[e*nsh+k]];
            ip_visc[e*nip+ip] = s + kin_visc;
        }
    free(D); free(vol); free(nt);
    return rc;
}

/* vorticityFV1 (navier_stokes_tools.h:386-525): per element and corner co
 *   localvort = sum_sh (u(sh)[1] scv.global_grad(sh)[0] - u(sh)[0] scv.global_grad(sh)[1]) * scv.volume()
 * (global gradients at the SCV ip = the corner, ugcore FV1Geometry SCV, our spec), summed per vertex, divided by the vertex volume. */
int ora_fv1_vorticity(int elem, int64_t n_elem, int64_t n_node, const int32_t *conn, const double *coords, const double *u, double *vort)
{
    const RefElem *r = get_ref(elem);
    if (!r) return fail("ora_fv1_vorticity: unknown element type");
    const int dim = r->dim, nsh = r->nsh, nf = dim + 1;
    double *vol = calloc((size_t)n_node, sizeof *vol);
    memset(vort, 0, sizeof(double) * (size_t)n_node);
    int rc = 0;
    for (int64_t e = 0; e < n_elem && !rc; e++) {
        double xc[MAXSH*3], x[MAXSH][3]; Geom g;
        for (int k = 0; k < nsh; k++) for (int d = 0; d < 3; d++) { x[k][d] = d < dim ? coords[(int64_t)conn[e*nsh+k]*dim+d] : 0.0; if (d < dim) xc[k*dim+d] = x[k][d]; }
        if ((rc = geom_update(&g, elem, xc))) break;
        for (int co = 0; co < nsh; co++) {
            double N[MAXSH], lg[MAXSH][3], JT[3][3] = {{0}}, JTinv[3][3] = {{0}}, w = 0;
            lagrange_shapes(elem, r->corner[co], N, lg);
            for (int i = 0; i < dim; i++) for (int j = 0; j < dim; j++) { double s = 0; for (int k = 0; k < nsh; k++) s += lg[k][i] * x[k][j]; JT[i][j] = s; }
            if (!(fabs(mat_inverse(dim, JT, JTinv)) > 0)) { rc = fail("FV1Geometry: singular element Jacobian"); break; }
            for (int k = 0; k < nsh; k++) {
                double G[2];
                for (int j = 0; j < 2; j++) { double s = 0; for (int i = 0; i < dim; i++) s += JTinv[j][i] * lg[k][i]; G[j] = s; }
                w += u[(int64_t)conn[e*nsh+k]*nf+1] * G[0] - u[(int64_t)conn[e*nsh+k]*nf+0] * G[1];
            }
            vort[conn[e*nsh+co]] += w * g.vol[co];
            vol[conn[e*nsh+co]] += g.vol[co];
        }
    }
    for (int64_t a = 0; a < n_node; a++) if (vol[a] > 0) vort[a] /= vol[a];
    free(vol);
    return rc;
}

/* kineticEnergy (navier_stokes_tools.h:850-965) and cflNumber (:731-848) of a Crouzeix-Raviart velocity field:
 *   value_e = sum_s N^CR_s(barycentre) u_s,  E = sum_e vol_e |value_e|^2 / sum_e vol_e   (vol_e = sum of the CR SCV volumes);
 *   cfl = max_e max_{i<j} dt |(x_i - x_j) . value_e| / |x_i - x_j|^2,  x_s = SCV ip of side s.
 * u: FVCR dof vector (side*dim+d). out[0] = kinetic energy, out[1] = max CFL number. */
int ora_fvcr_diagnostics(int elem, int64_t n_elem, const int32_t *conn, const double *coords, const int32_t *es, const double *u,
                         double dt, double *out)
{
    if (elem != ORA_TRI && elem != ORA_TET) return fail("ora_fvcr_diagnostics: simplices only");
    const int dim = ora_elem_dim(elem), nco = ora_elem_nsh(elem), ns = ora_elem_nside(elem);
    double E = 0, V = 0, cfl = 0;
    for (int64_t e = 0; e < n_elem; e++) {
        double xc[MAXSH*3], N[6], lb[3] = {0,0,0}; CRGeom g;
        for (int k = 0; k < nco; k++) for (int d = 0; d < dim; d++) xc[k*dim+d] = coords[(int64_t)conn[e*nco+k]*dim+d];
        if (cr_geom_update(&g, elem, xc)) return -1;
        for (int d = 0; d < dim; d++) lb[d] = 1.0 / (dim + 1);          /* global_to_local(barycentre) of an affine simplex */
        cr_shapes(elem, lb, N, NULL);
        double val[3] = {0,0,0}, ve = 0;
        for (int s = 0; s < ns; s++) { for (int d = 0; d < dim; d++) val[d] += N[s] * u[(int64_t)es[e*ns+s]*dim+d]; ve += g.vol[s]; }
        for (int d = 0; d < dim; d++) E += ve * val[d] * val[d];
        V += ve;
        for (int i = 0; i < ns; i++) for (int j = i+1; j < ns; j++) {
            double sub[3], q = 0, dd = 0;
            for (int d = 0; d < dim; d++) { sub[d] = g.scv_xip[i][d] - g.scv_xip[j][d]; q += sub[d] * val[d]; dd += sub[d] * sub[d]; }
            const double l = dt * 1.0 / dd * fabs(q);
            if (l > cfl) cfl = l;
        }
    }
    out[0] = E / V; out[1] = cfl;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * DiscConstraintFVCR (fvcr/disc_constraint_fvcr.h:164-1198), the default configuration
 * init(u, bLinUpConvDefect = true, false, bLinPressureDefect = true, false, bAdaptive = false, bLimiter = false) (:254-300):
 * post-assembly correction of the DEFECT of the FVCR discretisation (adjust_defect :1149-1171 -> add_defect :770-1147)
 *   1. side gradients (:780-870): acGrad(side) = sum_elem vol_scv * [sum_sh u_sh,d0 grad_sh,d1] / sum_elem vol_scv
 *   2. per element (skipped if one of its sides lies in a zero-gradient subset, :321-327, :1022-1024):
 *      StdVel(ip) = sum_sh u_sh shape_sh(ip);  flux = s_a StdVel . n;  base = flux > 0 ? from : to   (:1103-1108)
 *      linear upwind: upwindVel_d1 = acGrad(base)_d1 . (x_ip - x_scv(base)); d(d1, from) += upwindVel flux, d(d1, to) -= (:1118-1128)
 *      linear pressure (:1058-1090, :1109-1133): pGrad = 1/|elem| sum_sides n_side * (boundary side ? p_e : (p_e + p_nb)/2),
 *      pressure = s_a pGrad . (x_ip - barycentre);  d(d1, from) += pressure n_d1,  d(d1, to) -= pressure n_d1
 * (hanging nodes / bAdaptive, the Jacobian variants and the limiter are out of scope). u: FVCR dof vector; zero_grad_side
 * [n_side] flags or NULL; the correction is ADDED to defect.
 * ---------------------------------------------------------------------------------------- */
int ora_fvcr_constraint_defect(int elem, int64_t n_elem, int64_t n_side, const int32_t *conn, const double *coords, const int32_t *es,
                               const double *u, double s_a, int lin_upwind, int lin_pressure, const uint8_t *zero_grad_side,
                               double *defect)
{
    if (elem != ORA_TRI && elem != ORA_TET) return fail("ora_fvcr_constraint_defect: simplices only");
    const int dim = ora_elem_dim(elem), nco = ora_elem_nsh(elem), ns = ora_elem_nside(elem), nip = elem == ORA_TRI ? 3 : 6;
    double *grad = calloc((size_t)n_side * 9, sizeof *grad), *vol = calloc((size_t)n_side, sizeof *vol);
    int64_t *nb = malloc(sizeof(int64_t) * (size_t)n_side * 2);
    for (int64_t i = 0; i < n_side * 2; i++) nb[i] = -1;
    int rc = 0;
    for (int64_t e = 0; e < n_elem && !rc; e++) {
        double xc[MAXSH*3]; CRGeom g;
        for (int k = 0; k < nco; k++) for (int d = 0; d < dim; d++) xc[k*dim+d] = coords[(int64_t)conn[e*nco+k]*dim+d];
        if ((rc = cr_geom_update(&g, elem, xc))) break;
        double gg[3][3] = {{0}};
        for (int d0 = 0; d0 < dim; d0++) for (int sh = 0; sh < ns; sh++) for (int d1 = 0; d1 < dim; d1++)
            gg[d0][d1] += u[(int64_t)es[e*ns+sh]*dim+d0] * g.G[0][sh][d1];
        for (int s = 0; s < ns; s++) {
            const int64_t sd = es[e*ns+s];
            for (int d0 = 0; d0 < dim; d0++) for (int d1 = 0; d1 < dim; d1++) grad[sd*9+d0*3+d1] += gg[d0][d1] * g.vol[s];
            vol[sd] += g.vol[s];
            if (nb[sd*2] < 0) nb[sd*2] = e; else nb[sd*2+1] = e;
        }
    }
    for (int64_t sd = 0; sd < n_side; sd++) if (vol[sd] > 0) for (int i = 0; i < 9; i++) grad[sd*9+i] /= vol[sd];
    const int64_t pbase = n_side * dim;
    for (int64_t e = 0; e < n_elem && !rc; e++) {
        int skip = 0;
        if (zero_grad_side) for (int s = 0; s < ns; s++) if (zero_grad_side[es[e*ns+s]]) skip = 1;
        if (skip) continue;
        double xc[MAXSH*3]; CRGeom g;
        for (int k = 0; k < nco; k++) for (int d = 0; d < dim; d++) xc[k*dim+d] = coords[(int64_t)conn[e*nco+k]*dim+d];
        if ((rc = cr_geom_update(&g, elem, xc))) break;
        double pg[3] = {0,0,0};
        if (lin_pressure) {
            const double pe = u[pbase+e];
            double ve = 0;
            for (int s = 0; s < ns; s++) {
                const int64_t sd = es[e*ns+s];
                const int64_t other = nb[sd*2] == e ? nb[sd*2+1] : nb[sd*2];
                const double pv = other < 0 ? pe : 0.5 * (pe + u[pbase+other]);
                for (int d = 0; d < dim; d++) pg[d] += g.scv_n[s][d] * pv;
                ve += g.vol[s];
            }
            for (int d = 0; d < dim; d++) pg[d] /= ve;
        }
        for (int ip = 0; ip < nip; ip++) {
            double sv[3] = {0,0,0};
            for (int sh = 0; sh < ns; sh++) for (int d = 0; d < dim; d++) sv[d] += u[(int64_t)es[e*ns+sh]*dim+d] * g.N[ip][sh];
            const double flux = s_a * vdot(sv, g.n[ip], dim);
            const int base = flux > 0 ? g.from[ip] : g.to[ip];
            double pressure = 0;
            if (lin_pressure) { double q = 0; for (int j = 0; j < dim; j++) q += pg[j] * (g.xip[ip][j] - g.bary[j]); pressure = s_a * q; }
            const int64_t sf = es[e*ns+g.from[ip]], st = es[e*ns+g.to[ip]], sb = es[e*ns+base];
            for (int d1 = 0; d1 < dim; d1++) {
                if (lin_upwind) {
                    double uv = 0;
                    for (int d2 = 0; d2 < dim; d2++) uv += grad[sb*9+d1*3+d2] * (g.xip[ip][d2] - g.scv_xip[base][d2]);
                    defect[sf*dim+d1] += uv * flux; defect[st*dim+d1] -= uv * flux;
                }
                if (lin_pressure) { defect[sf*dim+d1] += pressure * g.n[ip][d1]; defect[st*dim+d1] -= pressure * g.n[ip][d1]; }
            }
        }
    }
    free(grad); free(vol); free(nb);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * Global level: CSR pattern = full element coupling incl. explicit zeros (App. B-7), dof
 * numbering App. B-8, serial element loop with AddLocalMatrixToGlobal-style scatter.
 * ---------------------------------------------------------------------------------------- */
static int cmp_i32(const void *a, const void *b) { int32_t x = *(const int32_t*)a, y = *(const int32_t*)b; return (x>y)-(x<y); }

/* generic: entities (nodes / sides) -> adjacent-entity lists through elements */
static int64_t entity_csr(int64_t n_elem, int64_t n_ent, int per, const int32_t *conn,
                          int64_t **optr, int32_t **oidx)
{
    int64_t *cnt = calloc((size_t)n_ent+1, sizeof *cnt);
    for (int64_t e = 0; e < n_elem; e++) for (int k = 0; k < per; k++) cnt[conn[e*per+k]+1]++;
    for (int64_t i = 0; i < n_ent; i++) cnt[i+1] += cnt[i];
    int64_t *e2 = malloc(sizeof(int64_t)*(size_t)(cnt[n_ent] ? cnt[n_ent] : 1));
    int64_t *pos = malloc(sizeof(int64_t)*(size_t)(n_ent+1)); memcpy(pos, cnt, sizeof(int64_t)*(size_t)(n_ent+1));
    for (int64_t e = 0; e < n_elem; e++) for (int k = 0; k < per; k++) e2[pos[conn[e*per+k]]++] = e;
    int64_t *ptr = malloc(sizeof(int64_t)*(size_t)(n_ent+1)); ptr[0] = 0;
    int64_t cap = 16; int32_t *idx = malloc(sizeof(int32_t)*(size_t)cap); int64_t nn = 0;
    int32_t tmp[4096];
    for (int64_t i = 0; i < n_ent; i++) {
        int m = 0;
        for (int64_t q = cnt[i]; q < cnt[i+1]; q++) for (int k = 0; k < per; k++) { if (m < 4096) tmp[m++] = conn[e2[q]*per+k]; }
        qsort(tmp, (size_t)m, sizeof(int32_t), cmp_i32);
        int u = 0; for (int k = 0; k < m; k++) if (k == 0 || tmp[k] != tmp[k-1]) tmp[u++] = tmp[k];
        if (nn + u > cap) { while (nn + u > cap) cap *= 2; idx = realloc(idx, sizeof(int32_t)*(size_t)cap); }
        memcpy(idx+nn, tmp, sizeof(int32_t)*(size_t)u); nn += u; ptr[i+1] = nn;
    }
    free(cnt); free(e2); free(pos);
    *optr = ptr; *oidx = idx; return nn;
}

int64_t ora_fv1_csr(int elem, int64_t n_elem, int64_t n_node, const int32_t *conn, int64_t *rowptr, int32_t *colind)
{
    int nsh = ora_elem_nsh(elem), nf = ora_elem_dim(elem)+1;
    int64_t *ptr; int32_t *idx;
    int64_t nb = entity_csr(n_elem, n_node, nsh, conn, &ptr, &idx);
    int64_t nnz = nb*nf*nf;
    if (rowptr) {
        int64_t pos = 0;
        for (int64_t i = 0; i < n_node; i++) for (int rf = 0; rf < nf; rf++) {
            rowptr[i*nf+rf] = pos;
            for (int64_t q = ptr[i]; q < ptr[i+1]; q++) for (int cf = 0; cf < nf; cf++) colind[pos++] = idx[q]*nf+cf;
        }
        rowptr[n_node*nf] = pos;
    }
    free(ptr); free(idx);
    return nnz;
}

int64_t ora_fvcr_csr(int elem, int64_t n_elem, int64_t n_side, const int32_t *es, int64_t *rowptr, int32_t *colind)
{
    int ns = ora_elem_nside(elem), dim = ora_elem_dim(elem);
    int64_t *ptr; int32_t *idx;
    int64_t nb = entity_csr(n_elem, n_side, ns, es, &ptr, &idx);
    /* side -> adjacent elements */
    int64_t *scnt = calloc((size_t)n_side+1, sizeof *scnt);
    for (int64_t e = 0; e < n_elem; e++) for (int k = 0; k < ns; k++) scnt[es[e*ns+k]+1]++;
    for (int64_t i = 0; i < n_side; i++) scnt[i+1] += scnt[i];
    int64_t *s2e = malloc(sizeof(int64_t)*(size_t)(scnt[n_side]+1)), *pos = malloc(sizeof(int64_t)*(size_t)(n_side+1));
    memcpy(pos, scnt, sizeof(int64_t)*(size_t)(n_side+1));
    for (int64_t e = 0; e < n_elem; e++) for (int k = 0; k < ns; k++) s2e[pos[es[e*ns+k]]++] = e;   /* ascending e */
    int64_t nnz = nb*dim*dim + scnt[n_side]*dim /*vel rows x p cols*/ + n_elem*(ns*dim+1);
    if (rowptr) {
        int64_t q = 0, pbase = n_side*dim;
        for (int64_t s = 0; s < n_side; s++) for (int d = 0; d < dim; d++) {
            rowptr[s*dim+d] = q;
            for (int64_t k = ptr[s]; k < ptr[s+1]; k++) for (int d2 = 0; d2 < dim; d2++) colind[q++] = idx[k]*dim+d2;
            for (int64_t k = scnt[s]; k < scnt[s+1]; k++) colind[q++] = (int32_t)(pbase + s2e[k]);
        }
        for (int64_t e = 0; e < n_elem; e++) {
            rowptr[pbase+e] = q;
            int32_t tmp[6]; for (int k = 0; k < ns; k++) tmp[k] = es[e*ns+k];
            qsort(tmp, (size_t)ns, sizeof(int32_t), cmp_i32);
            for (int k = 0; k < ns; k++) for (int d2 = 0; d2 < dim; d2++) colind[q++] = tmp[k]*dim+d2;
            colind[q++] = (int32_t)(pbase+e);
        }
        rowptr[pbase+n_elem] = q;
    }
    free(ptr); free(idx); free(scnt); free(s2e); free(pos);
    return nnz;
}

static inline int64_t csr_find(const int64_t *rowptr, const int32_t *colind, int64_t row, int32_t col)
{
    int64_t lo = rowptr[row], hi = rowptr[row+1]-1;
    while (lo <= hi) { int64_t mid = (lo+hi)>>1; if (colind[mid] == col) return mid; if (colind[mid] < col) lo = mid+1; else hi = mid-1; }
    return -1;
}

/* greedy element colouring (conflict = shared node/side) for the threaded baseline */
static int colour_elements(int64_t n_elem, int64_t n_ent, int per, const int32_t *conn, int32_t *colour)
{
    int64_t *mask = calloc((size_t)n_ent, sizeof *mask);     /* bitmask of colours used at entity */
    int ncol = 0;
    for (int64_t e = 0; e < n_elem; e++) {
        int64_t used = 0;
        for (int k = 0; k < per; k++) used |= mask[conn[e*per+k]];
        int c = 0; while (c < 63 && (used >> c) & 1) c++;
        colour[e] = c; if (c+1 > ncol) ncol = c+1;
        for (int k = 0; k < per; k++) mask[conn[e*per+k]] |= (int64_t)1 << c;
    }
    free(mask);
    return ncol;
}

static int assemble_one(const ora_params *p0, int64_t e, int64_t n_ent, const int32_t *conn,
                        const double *coords, const int32_t *es, const double *u, const double *sol0,
                        const double *sol1, const int64_t *rowptr, const int32_t *colind, int what,
                        double scale_a, double scale_m, double *values, double *defect)
{
    ora_params pe = *p0; pe.elem_index = e;                   /* the per-ip imports are read at this element's slice */
    const ora_params *p = &pe;
    int nco = ora_elem_nsh(p->elem), dim = ora_elem_dim(p->elem);
    double xc[MAXSH*3], ul[MAXL], s0[MAXL], s1[MAXL], Jl[MAXL*MAXL], dl[MAXL];
    int64_t gidx[MAXL]; int L;
    for (int k = 0; k < nco; k++) for (int d = 0; d < dim; d++) xc[k*dim+d] = coords[(int64_t)conn[e*nco+k]*dim+d];
    if (p->disc == ORA_DISC_FV1) {
        int nf = dim+1; L = nf*nco;
        for (int f = 0; f < nf; f++) for (int k = 0; k < nco; k++) {
            int64_t gi = (int64_t)conn[e*nco+k]*nf+f; gidx[f*nco+k] = gi; ul[f*nco+k] = u[gi];
            if (sol0) s0[f*nco+k] = sol0[gi];
            if (sol1) s1[f*nco+k] = sol1[gi];
        }
    } else {
        int ns = ora_elem_nside(p->elem); L = dim*ns+1;
        for (int d = 0; d < dim; d++) for (int k = 0; k < ns; k++) { int64_t gi = (int64_t)es[e*ns+k]*dim+d; gidx[d*ns+k] = gi; ul[d*ns+k] = u[gi]; }
        gidx[dim*ns] = n_ent*dim + e; ul[dim*ns] = u[gidx[dim*ns]];
    }
    /* stiffness and mass parts are evaluated separately (as ugcore does) and scaled */
    for (int pass = 0; pass < 2; pass++) {
        int w = pass == 0 ? (what & (ORA_JAC_A|ORA_DEF_A|ORA_RHS)) : (what & (ORA_JAC_M|ORA_DEF_M));
        if (!w) continue;
        double sc = pass == 0 ? scale_a : scale_m;
        memset(Jl, 0, sizeof(double)*(size_t)(L*L)); memset(dl, 0, sizeof(double)*(size_t)L);
        int rc;
        if (p->disc == ORA_DISC_FV1) {
            rc = ora_fv1_elem(p, xc, ul, sol0 ? s0 : NULL, sol1 ? s1 : NULL, w & ~ORA_RHS, Jl, dl);
            if (!rc && (w & ORA_RHS)) { double r[MAXL] = {0}; rc = ora_fv1_elem(p, xc, ul, NULL, NULL, ORA_RHS, NULL, r); for (int i = 0; i < L; i++) dl[i] -= r[i]; }
        } else {
            rc = ora_fvcr_elem(p, xc, ul, w & ~ORA_RHS, Jl, dl);
            if (!rc && (w & ORA_RHS)) { double r[MAXL] = {0}; rc = ora_fvcr_elem(p, xc, ul, ORA_RHS, NULL, r); for (int i = 0; i < L; i++) dl[i] -= r[i]; }
        }
        if (rc) return rc;
        if (w & (ORA_JAC_A|ORA_JAC_M)) {
            for (int i = 0; i < L; i++) for (int j = 0; j < L; j++) {
                if (pass == 1 && i != j) continue;                  /* mass part is diagonal */
                int64_t q = csr_find(rowptr, colind, gidx[i], (int32_t)gidx[j]);
                if (q < 0) return fail("assemble: entry not in CSR pattern");
                values[q] += sc*Jl[i*L+j];
            }
        }
        if (w & (ORA_DEF_A|ORA_DEF_M|ORA_RHS)) for (int i = 0; i < L; i++) defect[gidx[i]] += sc*dl[i];
    }
    return 0;
}

int ora_assemble(const ora_params *p, int64_t n_elem, int64_t n_ent, const int32_t *conn,
                 const double *coords, const int32_t *es, const double *u, const double *sol0,
                 const double *sol1, const int64_t *rowptr, const int32_t *colind, int what,
                 double scale_a, double scale_m, double *values, double *defect, int nthreads)
{
    int rc = 0;
    if (p->disc == ORA_DISC_FVCR && !es) return fail("FVCR needs elem_sides");
    if (nthreads <= 1) {
        for (int64_t e = 0; e < n_elem && !rc; e++)
            rc = assemble_one(p, e, n_ent, conn, coords, es, u, sol0, sol1, rowptr, colind, what, scale_a, scale_m, values, defect);
        return rc;
    }
    int per = p->disc == ORA_DISC_FV1 ? ora_elem_nsh(p->elem) : ora_elem_nside(p->elem);
    const int32_t *cc = p->disc == ORA_DISC_FV1 ? conn : es;
    int32_t *colour = malloc(sizeof(int32_t)*(size_t)n_elem);
    int ncol = colour_elements(n_elem, n_ent, per, cc, colour);
    /* bucket by colour */
    int64_t *cptr = calloc((size_t)ncol+1, sizeof *cptr), *order = malloc(sizeof(int64_t)*(size_t)n_elem);
    for (int64_t e = 0; e < n_elem; e++) cptr[colour[e]+1]++;
    for (int c = 0; c < ncol; c++) cptr[c+1] += cptr[c];
    { int64_t *pos = malloc(sizeof(int64_t)*(size_t)(ncol+1)); memcpy(pos, cptr, sizeof(int64_t)*(size_t)(ncol+1));
      for (int64_t e = 0; e < n_elem; e++) order[pos[colour[e]]++] = e; free(pos); }
    for (int c = 0; c < ncol; c++) {
#pragma omp parallel for num_threads(nthreads) schedule(static)
        for (int64_t q = cptr[c]; q < cptr[c+1]; q++) {
            int r = assemble_one(p, order[q], n_ent, conn, coords, es, u, sol0, sol1, rowptr, colind, what, scale_a, scale_m, values, defect);
            if (r) {
#pragma omp atomic write
                rc = r;
            }
        }
    }
    free(colour); free(cptr); free(order);
    return rc;
}

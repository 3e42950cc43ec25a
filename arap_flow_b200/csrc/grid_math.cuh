// grid_math.cuh -- per-pixel closed forms of the ARAP energy derivatives, specialised for
// UrShape == pixel grid (d_ij = u_i - u_j is a signed unit axis vector), in exactly the operation
// order of the arithmetic contract (restated independently, for general d, in
// oracle/arap_oracle.c).  Reference: arap_plan.t:13-23 (energy), ARAP/API/src/o.t:2129-2172 (J^T F +
// diagonal), o.t:2029-2089 (J^T J p), o.t:2375-2385 (cost); closed forms SURVEY.md 8(a) a-3 / a-4.
//
// Neighbour n: 0 = +x (d = (-1, 0)), 1 = -x (d = (1, 0)), 2 = +y (d = (0, -1)), 3 = -y (d = (0, 1)).
// With such d every product with d is exact, so e.g. R'(a) d collapses to a signed pick of (sin, cos):
//   n   R(a) d        R'(a) d
//   0   (-c, -s)      ( s, -c)
//   1   ( c,  s)      (-s,  c)
//   2   ( s, -c)      ( c,  s)
//   3   (-s,  c)      (-c, -s)
#pragma once
#include "common.cuh"

namespace arapb200 {

// ------------------------------------------------------------------------------------------ J^T J p
// Tile entry of a pixel j for this phase: (p_x, p_y, sin(a_j) * p_a, cos(a_j) * p_a).
struct JtjAcc {
    float sd0, sd1; // sum_j (pX_i - pX_j)
    float nb0, nb1; // sum_j (R'(a_j) d_ij) * pa_j
    float dd, dc;   // sum_j d.dp, sum_j d x dp
    float Sx, Sy;   // sum_j d_ij
    float nd;       // sum_j |d_ij|^2  (= number of valid neighbours)
};

__device__ __forceinline__ void jtj_zero(JtjAcc& a)
{
    a.sd0 = a.sd1 = a.nb0 = a.nb1 = a.dd = a.dc = a.Sx = a.Sy = a.nd = 0.0f;
}

template <int N>
__device__ __forceinline__ void jtj_nb(JtjAcc& a, float px, float py, const float4 Pj)
{
    const float dp0 = px - Pj.x, dp1 = py - Pj.y;
    a.sd0 = a.sd0 + dp0;
    a.sd1 = a.sd1 + dp1;
    if (N == 0) {
        a.nb0 = a.nb0 + Pj.z;
        a.nb1 = a.nb1 + (-Pj.w);
        a.dd = a.dd + (-dp0);
        a.dc = a.dc + (-dp1);
        a.Sx = a.Sx + (-1.0f);
    } else if (N == 1) {
        a.nb0 = a.nb0 + (-Pj.z);
        a.nb1 = a.nb1 + Pj.w;
        a.dd = a.dd + dp0;
        a.dc = a.dc + dp1;
        a.Sx = a.Sx + 1.0f;
    } else if (N == 2) {
        a.nb0 = a.nb0 + Pj.w;
        a.nb1 = a.nb1 + Pj.z;
        a.dd = a.dd + (-dp1);
        a.dc = a.dc + dp0;
        a.Sy = a.Sy + (-1.0f);
    } else {
        a.nb0 = a.nb0 + (-Pj.w);
        a.nb1 = a.nb1 + (-Pj.z);
        a.dd = a.dd + dp1;
        a.dc = a.dc + (-dp0);
        a.Sy = a.Sy + 1.0f;
    }
    a.nd = a.nd + 1.0f;
}

// Masked variant for strips that touch the object boundary: m = 1.0f for a valid neighbour, 0.0f otherwise.
// fmaf(1, x, acc) == acc + x and fmaf(0, x, acc) == acc exactly (x finite), so this is bit-identical to skipping
// invalid neighbours; Sx, Sy and nd are accumulated from the masks themselves (exact small integers).
template <int N>
__device__ __forceinline__ void jtj_nb_masked(JtjAcc& a, float px, float py, const float4 Pj, float m)
{
    const float dp0 = px - Pj.x, dp1 = py - Pj.y;
    a.sd0 = fmaf(m, dp0, a.sd0);
    a.sd1 = fmaf(m, dp1, a.sd1);
    if (N == 0) {
        a.nb0 = fmaf(m, Pj.z, a.nb0);
        a.nb1 = fmaf(-m, Pj.w, a.nb1);
        a.dd = fmaf(-m, dp0, a.dd);
        a.dc = fmaf(-m, dp1, a.dc);
        a.Sx = a.Sx - m;
    } else if (N == 1) {
        a.nb0 = fmaf(-m, Pj.z, a.nb0);
        a.nb1 = fmaf(m, Pj.w, a.nb1);
        a.dd = fmaf(m, dp0, a.dd);
        a.dc = fmaf(m, dp1, a.dc);
        a.Sx = a.Sx + m;
    } else if (N == 2) {
        a.nb0 = fmaf(m, Pj.w, a.nb0);
        a.nb1 = fmaf(m, Pj.z, a.nb1);
        a.dd = fmaf(-m, dp1, a.dd);
        a.dc = fmaf(m, dp0, a.dc);
        a.Sy = a.Sy - m;
    } else {
        a.nb0 = fmaf(-m, Pj.w, a.nb0);
        a.nb1 = fmaf(-m, Pj.z, a.nb1);
        a.dd = fmaf(m, dp1, a.dd);
        a.dc = fmaf(-m, dp0, a.dc);
        a.Sy = a.Sy + m;
    }
    a.nd = a.nd + m;
}

__device__ __forceinline__ void jtj_finish(const JtjAcc& a, float ci, float si, float px, float py, float pa,
                                           bool fit, float wr2, float wf2, float& q0, float& q1, float& qa)
{
    const float E0 = (-(si * a.Sx)) - ci * a.Sy;
    const float E1 = ci * a.Sx - si * a.Sy;
    const float own0 = E0 * pa, own1 = E1 * pa;
    float t0 = (a.sd0 + a.sd0) - own0;
    t0 = t0 - a.nb0;
    float t1 = (a.sd1 + a.sd1) - own1;
    t1 = t1 - a.nb1;
    q0 = wr2 * t0;
    q1 = wr2 * t1;
    const float rdp = fmaf(ci, a.dc, -(si * a.dd));
    qa = wr2 * fmaf(a.nd, pa, -rdp);
    if (fit) {
        q0 = fmaf(wf2, px, q0);
        q1 = fmaf(wf2, py, q1);
    }
}

// ------------------------------------------------------------------------------------------ J^T F
// Tile entry of a pixel j for this phase: (X_x, X_y, cos a_j, sin a_j).
struct JtfAcc {
    float gx0, gx1, ga, nd, nv;
};
__device__ __forceinline__ void jtf_zero(JtfAcc& a) { a.gx0 = a.gx1 = a.ga = a.nd = a.nv = 0.0f; }

// R(a) d for neighbour N
template <int N>
__device__ __forceinline__ void rot_d(float c, float s, float& r0, float& r1)
{
    if (N == 0) { r0 = -c; r1 = -s; }
    else if (N == 1) { r0 = c; r1 = s; }
    else if (N == 2) { r0 = s; r1 = -c; }
    else { r0 = -s; r1 = c; }
}
// R'(a) d for neighbour N
template <int N>
__device__ __forceinline__ void drot_d(float c, float s, float& q0, float& q1)
{
    if (N == 0) { q0 = s; q1 = -c; }
    else if (N == 1) { q0 = -s; q1 = c; }
    else if (N == 2) { q0 = c; q1 = s; }
    else { q0 = -c; q1 = -s; }
}

template <int N>
__device__ __forceinline__ void jtf_nb(JtfAcc& a, float X0, float X1, float ci, float si, const float4 Ej)
{
    const float dX0 = X0 - Ej.x, dX1 = X1 - Ej.y;
    float Ri0, Ri1, Rj0, Rj1, Q0, Q1;
    rot_d<N>(ci, si, Ri0, Ri1);
    rot_d<N>(Ej.z, Ej.w, Rj0, Rj1);
    drot_d<N>(ci, si, Q0, Q1);
    const float e0 = dX0 - Ri0, e1 = dX1 - Ri1;
    float t0 = (dX0 + dX0) - Ri0;
    t0 = t0 - Rj0;
    float t1 = (dX1 + dX1) - Ri1;
    t1 = t1 - Rj1;
    a.gx0 = a.gx0 + t0;
    a.gx1 = a.gx1 + t1;
    a.ga = a.ga + fmaf(Q1, e1, Q0 * e0);
    a.nd = a.nd + 1.0f;
    a.nv = a.nv + 1.0f;
}

// g = J^T F (3 comps) and the two distinct diagonal entries
__device__ __forceinline__ void jtf_finish(const JtfAcc& a, float X0, float X1, bool fit, float C0, float C1,
                                           float wr2, float wf2, float& g0, float& g1, float& ga, float& DX,
                                           float& DA)
{
    g0 = wr2 * a.gx0;
    g1 = wr2 * a.gx1;
    DX = (wr2 + wr2) * a.nv;
    if (fit) {
        g0 = fmaf(wf2, X0 - C0, g0);
        g1 = fmaf(wf2, X1 - C1, g1);
        DX = DX + wf2;
    }
    ga = -(wr2 * a.ga);
    DA = wr2 * a.nd;
}

// ------------------------------------------------------------------------------------------ cost
template <int N>
__device__ __forceinline__ float cost_nb(float acc, float X0, float X1, float ci, float si, const float4 Ej,
                                         float wr)
{
    const float dX0 = X0 - Ej.x, dX1 = X1 - Ej.y;
    float Ri0, Ri1;
    rot_d<N>(ci, si, Ri0, Ri1);
    const float e0 = dX0 - Ri0, e1 = dX1 - Ri1;
    const float w0 = wr * e0, w1 = wr * e1;
    acc = fmaf(w0, w0, acc);
    acc = fmaf(w1, w1, acc);
    return acc;
}
__device__ __forceinline__ float cost_fit(float acc, float X0, float X1, float C0, float C1, float wf)
{
    const float f0 = wf * (X0 - C0), f1 = wf * (X1 - C1);
    acc = fmaf(f0, f0, acc);
    acc = fmaf(f1, f1, acc);
    return acc;
}

// ------------------------------------------------------------------------------------------ general UrShape
// The same three derived functions for an arbitrary UrShape image: d = u_i - u_j is a runtime vector.  Operation order
// of oracle/arap_oracle.c (jtf_pixel, jtj_pixel, cost_pixel); with d a signed unit axis vector every expression below
// reduces bit for bit to the specialised forms above.  Used by the streaming back-end's *_gen kernels only.
__device__ __forceinline__ void jtj_nb_gen(JtjAcc& a, float px, float py, float pj0, float pj1, float paj, float cj,
                                           float sj, float dx, float dy)
{
    const float dp0 = px - pj0, dp1 = py - pj1;
    a.sd0 = a.sd0 + dp0;
    a.sd1 = a.sd1 + dp1;
    const float Qj0 = (-(sj * dx)) - cj * dy, Qj1 = cj * dx - sj * dy; // R'(a_j) d
    a.nb0 = a.nb0 + Qj0 * paj;
    a.nb1 = a.nb1 + Qj1 * paj;
    a.dd = a.dd + (dx * dp0 + dy * dp1);
    a.dc = a.dc + (dx * dp1 - dy * dp0);
    a.Sx = a.Sx + dx;
    a.Sy = a.Sy + dy;
    a.nd = a.nd + (dx * dx + dy * dy); // contract C4
}

__device__ __forceinline__ void jtf_nb_gen(JtfAcc& a, float X0, float X1, float ci, float si, float Xj0, float Xj1,
                                           float cj, float sj, float dx, float dy)
{
    const float dX0 = X0 - Xj0, dX1 = X1 - Xj1;
    const float Ri0 = ci * dx - si * dy, Ri1 = si * dx + ci * dy;
    const float Rj0 = cj * dx - sj * dy, Rj1 = sj * dx + cj * dy;
    const float e0 = dX0 - Ri0, e1 = dX1 - Ri1;
    float t0 = (dX0 + dX0) - Ri0;
    t0 = t0 - Rj0;
    float t1 = (dX1 + dX1) - Ri1;
    t1 = t1 - Rj1;
    a.gx0 = a.gx0 + t0;
    a.gx1 = a.gx1 + t1;
    const float Q0 = (-(si * dx)) - ci * dy, Q1 = ci * dx - si * dy; // R'(a_i) d
    a.ga = a.ga + fmaf(Q1, e1, Q0 * e0);
    a.nd = a.nd + (dx * dx + dy * dy);
    a.nv = a.nv + 1.0f;
}

__device__ __forceinline__ float cost_nb_gen(float acc, float X0, float X1, float ci, float si, float Xj0, float Xj1,
                                             float dx, float dy, float wr)
{
    const float dX0 = X0 - Xj0, dX1 = X1 - Xj1;
    const float Ri0 = ci * dx - si * dy, Ri1 = si * dx + ci * dy;
    const float e0 = dX0 - Ri0, e1 = dX1 - Ri1;
    const float w0 = wr * e0, w1 = wr * e1;
    acc = fmaf(w0, w0, acc);
    acc = fmaf(w1, w1, acc);
    return acc;
}

} // namespace arapb200

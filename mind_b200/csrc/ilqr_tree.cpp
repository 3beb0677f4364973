// Tree iLQR of MIND's trajectory-tree optimiser, native host code (sequential passes over tiny 6x6 /
// 2x2 systems -- host C++ is the right place for it; the cost fields it consumes come from csrc/cost_field.cu).
// Follows the reference's algorithm step for step so that the iterates coincide:
//   planners/ilqr/solver.py:80-167   fit: Levenberg-Marquardt schedule, 10 backtracking steps alpha = 1.1^-(k^2)
//   solver.py:261-329                forward rollout over the tree (Jacobians evaluated at the NEW state, as the reference does)
//   solver.py:332-373                recursive backward pass, children's value expansions summed into the parent
//   solver.py:375-421                Q expansion with the regulariser on V_xx inside Q_ux / Q_uu only
//   solver.py:202-240, 242-254      line search roll-out, trajectory cost (left-to-right Python sum)
//   planners/ilqr/cost.py:326-446    TreeCost: per node [PotentialField, StatePotential, StateConstraint] + [ControlPotential]
//   planners/ilqr/potential.py       the four potentials; PotentialField = bi-quadratic B-spline patch on a 3x3 neighbourhood
//                                    smoothed by 2x2 / 1x2 means, cell picked with round-half-even, zero padding at the border
//   planners/mind/trajectory_tree.py:153-177   6-state kinematic bicycle model (x, y, v, heading, a, steer; inputs da, dsteer)
// numpy's pairwise summation (J_opt = L.sum()) and its sequential 2x2 mean are reproduced because accept / converge
// decisions compare sums.
#include "../../include/mind_b200.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

namespace {

constexpr int NX = 6, NU = 2;
thread_local char g_ierr[256] = "";

double np_sum(const double* a, int n) {            // numpy's pairwise add.reduce on a contiguous fp64 vector
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += a[i];
        return r;
    }
    if (n <= 128) {
        double r[8];
        for (int k = 0; k < 8; ++k) r[k] = a[k];
        int i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] += a[i + k];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return np_sum(a, n2) + np_sum(a + n2, n - n2);
}

struct Problem {
    const MindIlqrTree* p;
    int n;
    std::vector<std::vector<int>> kids;      // children in creation order; kids_root = children of the root state
    std::vector<int> kids_root;

    // ---- dynamics (trajectory_tree.py:153-177) ----
    void f(const double* x, const double* u, double* o) const {
        const double dt = p->dt, wb = p->wheelbase;
        o[0] = x[0] + x[2] * std::cos(x[3]) * dt;
        o[1] = x[1] + x[2] * std::sin(x[3]) * dt;
        o[2] = x[2] + x[4] * dt;
        o[3] = x[3] + x[2] / wb * std::tan(x[5]) * dt;
        o[4] = x[4] + u[0] * dt;
        o[5] = x[5] + u[1] * dt;
    }
    void fx(const double* x, double F[NX][NX]) const {
        const double dt = p->dt, wb = p->wheelbase;
        std::memset(F, 0, sizeof(double) * NX * NX);
        for (int i = 0; i < NX; ++i) F[i][i] = 1.0;
        const double c = std::cos(x[3]), s = std::sin(x[3]), t = std::tan(x[5]), c5 = std::cos(x[5]);
        F[0][2] = c * dt;           F[0][3] = (x[2] * (-s)) * dt;
        F[1][2] = s * dt;           F[1][3] = (x[2] * c) * dt;
        F[2][4] = dt;
        F[3][2] = ((1.0 / wb) * t) * dt;
        F[3][5] = ((x[2] / wb) * (1.0 / (c5 * c5))) * dt;
    }
    void fu(double G[NX][NU]) const {
        std::memset(G, 0, sizeof(double) * NX * NU);
        G[4][0] = p->dt;
        G[5][1] = p->dt;
    }

    // ---- PotentialField (potential.py:62-264) ----
    struct Patch { double S[3][3], u, v; };
    Patch patch(int node, const double* x) const {
        const int W = p->gx, H = p->gy;
        const double* F = p->fields + (size_t)node * W * H;
        auto at = [&](int r, int c) { return F[(size_t)r * W + c]; };
        long xi = std::lrint(std::nearbyint((x[0] - p->field_offset[0]) / p->res));    // Python round(): half to even
        long yi = std::lrint(std::nearbyint((x[1] - p->field_offset[1]) / p->res));
        xi = xi < 0 ? 0 : (xi > W - 1 ? W - 1 : xi);
        yi = yi < 0 ? 0 : (yi > H - 1 ? H - 1 : yi);
        double L[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        const int X = (int)xi, Y = (int)yi;
        if (X == 0 && Y == 0) {
            for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) L[1 + a][1 + b] = at(a, b);
        } else if (X == 0 && Y == H - 1) {
            for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) L[1 + a][b] = at(H - 2 + a, b);
        } else if (X == W - 1 && Y == 0) {
            for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) L[a][1 + b] = at(a, W - 2 + b);
        } else if (X == W - 1 && Y == H - 1) {
            for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) L[a][b] = at(H - 2 + a, W - 2 + b);
        } else if (X == 0) {
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 2; ++b) L[a][b] = at(Y - 1 + a, b);
        } else if (X == W - 1) {
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 2; ++b) L[a][1 + b] = at(Y - 1 + a, W - 2 + b);
        } else if (Y == 0) {
            for (int a = 0; a < 2; ++a) for (int b = 0; b < 3; ++b) L[a][b] = at(a, X - 1 + b);
        } else if (Y == H - 1) {
            for (int a = 0; a < 2; ++a) for (int b = 0; b < 3; ++b) L[1 + a][b] = at(H - 2 + a, X - 1 + b);
        } else {
            for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) L[a][b] = at(Y - 1 + a, X - 1 + b);
        }
        Patch q;
        q.S[0][0] = (((L[0][0] + L[0][1]) + L[1][0]) + L[1][1]) / 4.0;
        q.S[0][2] = (((L[0][1] + L[0][2]) + L[1][1]) + L[1][2]) / 4.0;
        q.S[2][0] = (((L[1][0] + L[1][1]) + L[2][0]) + L[2][1]) / 4.0;
        q.S[2][2] = (((L[1][1] + L[1][2]) + L[2][1]) + L[2][2]) / 4.0;
        q.S[0][1] = (L[0][1] + L[1][1]) / 2.0;
        q.S[1][0] = (L[1][0] + L[1][1]) / 2.0;
        q.S[1][2] = (L[1][1] + L[1][2]) / 2.0;
        q.S[2][1] = (L[1][1] + L[2][1]) / 2.0;
        q.S[1][1] = L[1][1];
        q.u = (x[0] - p->xs_grid[X]) / p->res + 0.5;
        q.v = (x[1] - p->ys_grid[Y]) / p->res + 0.5;
        return q;
    }
    static double sq(double a) { return a * a; }
    double field_value(const Patch& q) const {
        const double u = q.u, v = q.v;
        const double (*g)[3] = q.S;
        return sq(1 - u) * sq(1 - v) * g[0][0] + sq(1 - u) * 2.0 * (1 - v) * v * g[1][0] + sq(1 - u) * sq(v) * g[2][0] +
               2.0 * (1 - u) * u * sq(1 - v) * g[0][1] + 2.0 * (1 - u) * u * 2.0 * (1 - v) * v * g[1][1] +
               2.0 * (1 - u) * u * sq(v) * g[2][1] + sq(u) * sq(1 - v) * g[0][2] + sq(u) * 2.0 * (1 - v) * v * g[1][2] +
               sq(u) * sq(v) * g[2][2];
    }
    void field_grad(const Patch& q, double* gx, double* gy) const {
        const double u = q.u, v = q.v, r = p->res;
        const double (*g)[3] = q.S;
        *gx = 1.0 / r * ((-2.0 + 2.0 * u) * sq(1.0 - v) * g[0][0] + (-2.0 + 2.0 * u) * 2.0 * (1.0 - v) * v * g[1][0] +
                         (-2.0 + 2.0 * u) * sq(v) * g[2][0] + 2.0 * (1.0 - 2.0 * u) * sq(1.0 - v) * g[0][1] +
                         2.0 * (1.0 - 2.0 * u) * 2.0 * (1.0 - v) * v * g[1][1] + 2.0 * (1.0 - 2.0 * u) * sq(v) * g[2][1] +
                         u * 2.0 * sq(1.0 - v) * g[0][2] + u * 2.0 * 2.0 * (1.0 - v) * v * g[1][2] + u * 2.0 * sq(v) * g[2][2]);
        *gy = 1.0 / r * (sq(1.0 - u) * (-2.0 + 2.0 * v) * g[0][0] + sq(1.0 - u) * 2.0 * (1.0 - 2.0 * v) * g[1][0] +
                         sq(1.0 - u) * 2.0 * v * g[2][0] + 2.0 * (1.0 - u) * u * (-2.0 + 2.0 * v) * g[0][1] +
                         2.0 * (1.0 - u) * u * 2.0 * (1.0 - 2.0 * v) * g[1][1] + 2.0 * (1.0 - u) * u * 2.0 * v * g[2][1] +
                         sq(u) * (-2.0 + 2.0 * v) * g[0][2] + sq(u) * 2.0 * (1.0 - 2.0 * v) * g[1][2] + sq(u) * 2.0 * v * g[2][2]);
    }
    void field_hess(const Patch& q, double* hxx, double* hxy, double* hyy) const {
        const double u = q.u, v = q.v, r2 = sq(p->res);
        const double (*g)[3] = q.S;
        *hxx = 1.0 / r2 * (2.0 * sq(1.0 - v) * g[0][0] + 2.0 * (1.0 - v) * 2.0 * v * g[1][0] + 2.0 * sq(v) * g[2][0] +
                           -4.0 * sq(1.0 - v) * g[0][1] + -4.0 * (1.0 - v) * 2.0 * v * g[1][1] + -4.0 * sq(v) * g[2][1] +
                           2.0 * sq(1.0 - v) * g[0][2] + 2.0 * (1.0 - v) * 2.0 * v * g[1][2] + 2.0 * sq(v) * g[2][2]);
        *hyy = 1.0 / r2 * (2.0 * sq(1.0 - u) * g[0][0] + -4.0 * sq(1.0 - u) * g[1][0] + 2.0 * sq(1.0 - u) * g[2][0] +
                           2.0 * (1.0 - u) * 2.0 * u * g[0][1] + -4.0 * (1.0 - u) * 2.0 * u * g[1][1] +
                           2.0 * (1.0 - u) * 2.0 * u * g[2][1] + 2.0 * sq(u) * g[0][2] + -4.0 * sq(u) * g[1][2] +
                           2.0 * sq(u) * g[2][2]);
        *hxy = 1.0 / r2 * ((-2.0 + 2.0 * u) * (-2.0 + 2.0 * v) * g[0][0] + (-2.0 + 2.0 * u) * 2.0 * (1.0 - 2.0 * v) * g[1][0] +
                           (-2.0 + 2.0 * u) * 2.0 * v * g[2][0] + 2.0 * (1.0 - 2.0 * u) * (-2.0 + 2.0 * v) * g[0][1] +
                           2.0 * (1.0 - 2.0 * u) * 2.0 * (1.0 - 2.0 * v) * g[1][1] + 2.0 * (1.0 - 2.0 * u) * 2.0 * v * g[2][1] +
                           2.0 * u * (-2.0 + 2.0 * v) * g[0][2] + 2.0 * u * 2.0 * (1.0 - 2.0 * v) * g[1][2] +
                           2.0 * u * 2.0 * v * g[2][2]);
    }

    // ---- node cost (cost.py:341-446 over potential.py) ----
    static double quad_form(const double* W, const double* d, int n) {       // d^T W d, numpy order: (d.W).d
        double acc = 0.0;
        for (int j = 0; j < n; ++j) {
            double col = 0.0;
            for (int i = 0; i < n; ++i) col += d[i] * W[i * n + j];
            acc += col * d[j];
        }
        return acc;
    }
    double l(int node, const double* x, const double* u) const {
        double cost = 0.0;
        cost += field_value(patch(node, x));
        double d[NX];
        for (int i = 0; i < NX; ++i) d[i] = x[i] - p->des_state[(size_t)node * NX + i];
        cost += quad_form(p->w_state + (size_t)node * NX * NX, d, NX);
        for (int i = 0; i < NX; ++i) d[i] = std::fmax(x[i] - p->upper[i], 0.0) + std::fmax(p->lower[i] - x[i], 0.0);
        cost += quad_form(p->w_con + (size_t)node * NX * NX, d, NX);
        cost += quad_form(p->w_ctrl + (size_t)node * NU * NU, u, NU);
        return cost;
    }
    void l_derivs(int node, const double* x, const double* u, double* lx, double* lu, double lxx[NX][NX], double luu[NU][NU]) const {
        const double* Ws = p->w_state + (size_t)node * NX * NX;
        const double* Wc = p->w_con + (size_t)node * NX * NX;
        const double* Wu = p->w_ctrl + (size_t)node * NU * NU;
        const Patch q = patch(node, x);
        double gx, gy, hxx, hxy, hyy;
        field_grad(q, &gx, &gy);
        field_hess(q, &hxx, &hxy, &hyy);
        for (int i = 0; i < NX; ++i) lx[i] = 0.0;
        std::memset(lxx, 0, sizeof(double) * NX * NX);
        lx[0] += gx; lx[1] += gy;
        lxx[0][0] += hxx; lxx[0][1] += hxy; lxx[1][0] += hxy; lxx[1][1] += hyy;
        double d[NX];
        for (int i = 0; i < NX; ++i) d[i] = x[i] - p->des_state[(size_t)node * NX + i];
        for (int i = 0; i < NX; ++i) {
            double s = 0.0;
            for (int j = 0; j < NX; ++j) s += Ws[i * NX + j] * d[j];
            lx[i] += 2.0 * s;
            for (int j = 0; j < NX; ++j) lxx[i][j] += 2.0 * Ws[i * NX + j];
        }
        for (int i = 0; i < NX; ++i) {
            if (x[i] > p->upper[i]) { lx[i] += 2.0 * Wc[i * NX + i] * (x[i] - p->upper[i]); lxx[i][i] += 2.0 * Wc[i * NX + i]; }
            else if (x[i] < p->lower[i]) { lx[i] += 2.0 * Wc[i * NX + i] * (x[i] - p->lower[i]); lxx[i][i] += 2.0 * Wc[i * NX + i]; }
        }
        for (int i = 0; i < NU; ++i) {
            double s = 0.0;
            for (int j = 0; j < NU; ++j) s += Wu[i * NU + j] * u[j];
            lu[i] = 2.0 * s;
            for (int j = 0; j < NU; ++j) luu[i][j] = 2.0 * Wu[i * NU + j];
        }
    }
};

// 2x2 solve with partial pivoting (what LAPACK gesv does for numpy.linalg.solve); false = singular (LinAlgError)
bool solve2(const double A[NU][NU], const double* b, int nrhs, double* x) {     // b, x: [2][nrhs] row-major
    double a[2][2] = {{A[0][0], A[0][1]}, {A[1][0], A[1][1]}};
    int r0 = 0, r1 = 1;
    if (std::fabs(a[1][0]) > std::fabs(a[0][0])) { r0 = 1; r1 = 0; }
    if (a[r0][0] == 0.0) return false;
    const double m = a[r1][0] / a[r0][0];
    const double u11 = a[r1][1] - m * a[r0][1];
    if (u11 == 0.0) return false;
    for (int c = 0; c < nrhs; ++c) {
        const double y0 = b[r0 * nrhs + c], y1 = b[r1 * nrhs + c] - m * y0;
        const double x1 = y1 / u11;
        x[1 * nrhs + c] = x1;
        x[0 * nrhs + c] = (y0 - a[r0][1] * x1) / a[r0][0];
    }
    return true;
}

struct Solver {
    Problem P;
    int n;
    std::vector<double> xs, us, Fx, Fu, L, Lx, Lu, Lxx, Luu, Vx, Vxx, k, K;
    double J_opt = 0.0, mu = 1.0, delta = 2.0;

    explicit Solver(const MindIlqrTree* p) {
        P.p = p; n = P.n = p->n_nodes;
        P.kids.assign(n, {});
        for (int i = 0; i < n; ++i) {
            if (p->parent[i] < 0) P.kids_root.push_back(i);
            else P.kids[p->parent[i]].push_back(i);
        }
        xs.assign((size_t)n * NX, 0); us.assign((size_t)n * NU, 0); Fx.assign((size_t)n * NX * NX, 0); Fu.assign((size_t)n * NX * NU, 0);
        L.assign(n, 0); Lx.assign((size_t)n * NX, 0); Lu.assign((size_t)n * NU, 0); Lxx.assign((size_t)n * NX * NX, 0);
        Luu.assign((size_t)n * NU * NU, 0); Vx.assign((size_t)n * NX, 0); Vxx.assign((size_t)n * NX * NX, 0);
        k.assign((size_t)n * NU, 0); K.assign((size_t)n * NU * NX, 0);
    }
    void expand(int i, const double* prev) {                                   // one node of the forward rollout
        double* x = &xs[(size_t)i * NX];
        const double* u = &us[(size_t)i * NU];
        P.f(prev, u, x);
        P.fx(x, reinterpret_cast<double (*)[NX]>(&Fx[(size_t)i * NX * NX]));
        P.fu(reinterpret_cast<double (*)[NU]>(&Fu[(size_t)i * NX * NU]));
        L[i] = P.l(i, x, u);
        P.l_derivs(i, x, u, &Lx[(size_t)i * NX], &Lu[(size_t)i * NU], reinterpret_cast<double (*)[NX]>(&Lxx[(size_t)i * NX * NX]),
                   reinterpret_cast<double (*)[NU]>(&Luu[(size_t)i * NU * NU]));
    }
    void forward_rollout() {
        std::vector<int> stack;
        for (int c : P.kids_root) { expand(c, P.p->x0); stack.push_back(c); }
        while (!stack.empty()) {
            const int par = stack.back(); stack.pop_back();
            for (int c : P.kids[par]) { expand(c, &xs[(size_t)par * NX]); if (!P.kids[c].empty()) stack.push_back(c); }
        }
        J_opt = np_sum(L.data(), n);
    }
    bool gains(int i) {                                                        // solver.py:345-373 + 375-421
        const double (*fx)[NX] = reinterpret_cast<const double (*)[NX]>(&Fx[(size_t)i * NX * NX]);
        const double (*fu)[NU] = reinterpret_cast<const double (*)[NU]>(&Fu[(size_t)i * NX * NU]);
        const double* vx = &Vx[(size_t)i * NX];
        const double (*vxx)[NX] = reinterpret_cast<const double (*)[NX]>(&Vxx[(size_t)i * NX * NX]);
        double Qx[NX], Qu[NU], Qxx[NX][NX], Qux[NU][NX], Quu[NU][NU];
        for (int a = 0; a < NX; ++a) { double s = 0; for (int b = 0; b < NX; ++b) s += fx[b][a] * vx[b]; Qx[a] = Lx[(size_t)i * NX + a] + s; }
        for (int a = 0; a < NU; ++a) { double s = 0; for (int b = 0; b < NX; ++b) s += fu[b][a] * vx[b]; Qu[a] = Lu[(size_t)i * NU + a] + s; }
        double T[NX][NX];                                                      // f_x^T V_xx
        for (int a = 0; a < NX; ++a) for (int b = 0; b < NX; ++b) { double s = 0; for (int c = 0; c < NX; ++c) s += fx[c][a] * vxx[c][b]; T[a][b] = s; }
        for (int a = 0; a < NX; ++a) for (int b = 0; b < NX; ++b) { double s = 0; for (int c = 0; c < NX; ++c) s += T[a][c] * fx[c][b]; Qxx[a][b] = Lxx[(size_t)i * NX * NX + a * NX + b] + s; }
        double Tu[NU][NX];                                                     // f_u^T (V_xx + mu I)
        for (int a = 0; a < NU; ++a) for (int b = 0; b < NX; ++b) { double s = 0; for (int c = 0; c < NX; ++c) s += fu[c][a] * (vxx[c][b] + (c == b ? mu : 0.0)); Tu[a][b] = s; }
        for (int a = 0; a < NU; ++a) for (int b = 0; b < NX; ++b) { double s = 0; for (int c = 0; c < NX; ++c) s += Tu[a][c] * fx[c][b]; Qux[a][b] = 0.0 + s; }
        for (int a = 0; a < NU; ++a) for (int b = 0; b < NU; ++b) { double s = 0; for (int c = 0; c < NX; ++c) s += Tu[a][c] * fu[c][b]; Quu[a][b] = Luu[(size_t)i * NU * NU + a * NU + b] + s; }
        double kk[NU], KK[NU][NX];
        if (!solve2(Quu, Qu, 1, kk)) return false;
        if (!solve2(Quu, &Qux[0][0], NX, &KK[0][0])) return false;
        for (int a = 0; a < NU; ++a) { kk[a] = -kk[a]; for (int b = 0; b < NX; ++b) KK[a][b] = -KK[a][b]; }
        for (int a = 0; a < NU; ++a) { k[(size_t)i * NU + a] = kk[a]; for (int b = 0; b < NX; ++b) K[(size_t)i * NU * NX + a * NX + b] = KK[a][b]; }
        // V_x = Q_x + (K^T Q_uu) k + K^T Q_u + Q_ux^T k ;  V_xx = Q_xx + (K^T Q_uu) K + K^T Q_ux + Q_ux^T K, symmetrised
        double M[NX][NU];                                                      // K^T Q_uu, formed first as numpy's chained dot does
        for (int a = 0; a < NX; ++a) for (int c = 0; c < NU; ++c) M[a][c] = KK[0][a] * Quu[0][c] + KK[1][a] * Quu[1][c];
        double* ovx = &Vx[(size_t)i * NX];
        double (*ovxx)[NX] = reinterpret_cast<double (*)[NX]>(&Vxx[(size_t)i * NX * NX]);
        double nv[NX], nvv[NX][NX];
        for (int a = 0; a < NX; ++a) {
            nv[a] = Qx[a] + (M[a][0] * kk[0] + M[a][1] * kk[1]);
            nv[a] += (KK[0][a] * Qu[0] + KK[1][a] * Qu[1]) + (Qux[0][a] * kk[0] + Qux[1][a] * kk[1]);
            for (int b = 0; b < NX; ++b) {
                nvv[a][b] = Qxx[a][b] + (M[a][0] * KK[0][b] + M[a][1] * KK[1][b]);
                nvv[a][b] += (KK[0][a] * Qux[0][b] + KK[1][a] * Qux[1][b]) + (Qux[0][a] * KK[0][b] + Qux[1][a] * KK[1][b]);
            }
        }
        for (int a = 0; a < NX; ++a) { ovx[a] = nv[a]; for (int b = 0; b < NX; ++b) ovxx[a][b] = 0.5 * (nvv[a][b] + nvv[b][a]); }
        return true;
    }
    bool backward(int key) {                                                   // key = -1: root state
        const std::vector<int>& ch = key < 0 ? P.kids_root : P.kids[key];
        for (int c : ch) {
            if (!backward(c)) return false;
            if (!gains(c)) return false;
            if (key >= 0) {
                for (int a = 0; a < NX; ++a) Vx[(size_t)key * NX + a] += Vx[(size_t)c * NX + a];
                for (int a = 0; a < NX * NX; ++a) Vxx[(size_t)key * NX * NX + a] += Vxx[(size_t)c * NX * NX + a];
            }
        }
        return true;
    }
    bool backward_pass() {
        std::fill(Vx.begin(), Vx.end(), 0.0); std::fill(Vxx.begin(), Vxx.end(), 0.0);
        std::fill(k.begin(), k.end(), 0.0); std::fill(K.begin(), K.end(), 0.0);
        return backward(-1);
    }
    void line_search(double alpha, std::vector<double>& xn, std::vector<double>& un) const {
        xn.assign((size_t)n * NX, 0.0); un.assign((size_t)n * NU, 0.0);
        std::vector<int> stack;
        auto step = [&](int c, const double* xpar_new, const double* xpar_old) {
            for (int a = 0; a < NU; ++a) {
                double fb = 0.0;
                if (xpar_old) for (int b = 0; b < NX; ++b) fb += K[(size_t)c * NU * NX + a * NX + b] * (xpar_new[b] - xpar_old[b]);
                un[(size_t)c * NU + a] = xpar_old ? us[(size_t)c * NU + a] + alpha * k[(size_t)c * NU + a] + fb
                                                  : us[(size_t)c * NU + a] + alpha * k[(size_t)c * NU + a];
            }
            P.f(xpar_new, &un[(size_t)c * NU], &xn[(size_t)c * NX]);
        };
        // solver.py:222-223 treats node 0 as the only child of the root state
        step(0, P.p->x0, nullptr);
        stack.push_back(0);
        while (!stack.empty()) {
            const int par = stack.back(); stack.pop_back();
            for (int c : P.kids[par]) { step(c, &xn[(size_t)par * NX], &xs[(size_t)par * NX]); if (!P.kids[c].empty()) stack.push_back(c); }
        }
    }
    double trajectory_cost(const std::vector<double>& xn, const std::vector<double>& un) const {
        double J = 0.0;                                                        // Python sum(): left to right from 0
        for (int i = 0; i < n; ++i) J += P.l(i, &xn[(size_t)i * NX], &un[(size_t)i * NU]);
        return J;
    }
    int fit(int max_iter) {
        const double mu_min = 1e-6, mu_max = 1e10, delta0 = 2.0, rel_tol = 1e-6;
        mu = 1.0; delta = delta0;
        double alphas[10];
        for (int a = 0; a < 10; ++a) alphas[a] = std::pow(1.1, -(double)(a * a));
        for (int i = 0; i < n * NU; ++i) us[i] = P.p->us_init[i];
        bool accepted = true;
        int it = 0;
        std::vector<double> xn, un;
        for (; it < max_iter; ++it) {
            if (accepted) { forward_rollout(); accepted = false; }
            if (!backward_pass()) continue;                                    // LinAlgError: next iteration, nothing changes (solver.py:152-156)
            bool converged = false;
            for (int a = 0; a < 10; ++a) {
                line_search(alphas[a], xn, un);
                const double Jn = trajectory_cost(xn, un);
                if (Jn < J_opt) {
                    if (std::fabs((J_opt - Jn) / J_opt) < rel_tol) converged = true;
                    accepted = true;
                    xs = xn; us = un;
                    delta = std::fmin(1.0, delta) / delta0;
                    mu *= delta;
                    if (mu <= mu_min) mu = 0.0;
                    break;
                }
            }
            if (converged) { ++it; break; }
            if (!accepted) {
                delta = std::fmax(1.0, delta) * delta0;
                mu = std::fmax(mu_min, mu * delta);
                if (mu >= mu_max) { ++it; break; }
            }
        }
        return it;
    }
};

}  // namespace

extern "C" const char* mind_ilqr_last_error(void) { return g_ierr; }

// diagnostic: value, gradient and Hessian of one node's PotentialField at a position (potential.py:71-104)
extern "C" int mind_debug_field_eval(const MindIlqrTree* p, int32_t node, double x, double y, double* out6) {
    if (!p || !out6 || node < 0 || node >= p->n_nodes || !p->fields || !p->xs_grid || !p->ys_grid || !p->field_offset || p->gx < 2 ||
        p->gy < 2 || !(p->res > 0)) {
        snprintf(g_ierr, sizeof g_ierr, "mind_debug_field_eval: bad argument");
        return 1;
    }
    Problem P;
    P.p = p; P.n = p->n_nodes;
    const double pos[NX] = {x, y, 0, 0, 0, 0};
    const Problem::Patch q = P.patch(node, pos);
    out6[0] = P.field_value(q);
    P.field_grad(q, &out6[1], &out6[2]);
    P.field_hess(q, &out6[3], &out6[4], &out6[5]);       // xx, xy, yy
    return 0;
}

extern "C" int mind_ilqr_tree_solve(const MindIlqrTree* p) {
    if (!p || p->n_nodes <= 0 || !p->parent || !p->x0 || !p->fields || !p->xs_grid || !p->ys_grid || !p->field_offset ||
        !p->w_state || !p->des_state || !p->w_con || !p->lower || !p->upper || !p->w_ctrl || !p->us_init || !p->xs_out || !p->us_out ||
        p->gx < 2 || p->gy < 2 || !(p->res > 0) || !(p->dt > 0) || !(p->wheelbase > 0)) {
        snprintf(g_ierr, sizeof g_ierr, "mind_ilqr_tree_solve: bad argument");
        return 1;
    }
    if (p->parent[0] != -1) { snprintf(g_ierr, sizeof g_ierr, "mind_ilqr_tree_solve: node 0 must be the child of the root state"); return 1; }
    for (int i = 1; i < p->n_nodes; ++i)
        if (p->parent[i] < 0 || p->parent[i] >= i) { snprintf(g_ierr, sizeof g_ierr, "mind_ilqr_tree_solve: parent[%d] = %d (must precede its child)", i, p->parent[i]); return 1; }
    Solver s(p);
    const int it = s.fit(p->max_iter > 0 ? p->max_iter : 100);
    std::memcpy(p->xs_out, s.xs.data(), sizeof(double) * s.xs.size());
    std::memcpy(p->us_out, s.us.data(), sizeof(double) * s.us.size());
    if (p->iterations) *p->iterations = it;
    if (p->cost) *p->cost = s.J_opt;
    return 0;
}

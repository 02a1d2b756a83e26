// Host-side construction of the DSP constant tables, mirroring the reference's own derivations INCLUDING the
// precision each step is evaluated in (float32 torch tensors, float32 numpy scalars under NEP 50, complex128 numpy):
//   RADAE.__init__           radae/radae.py:172-219   w, Winv, Wfwd, P, Pend, p, pend, pilot_gain, EOO frame
//   acquisition.__init__     radae/dsp.py:153-176     fcoarse, p_w
//   receiver_one.__init__    radae/dsp.py:401-412     Pmat (plain transpose, not conjugate), exp(-j w_c a)
//   complex_bpf.__init__     radae/dsp.py:40-61 with the float32 arguments radae_rxe.py:104-109 passes
// Checked against oracle/dsp.py (itself pinned against the reference) by tests/test_tables.py through
// rade_b200_debug_tables().
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <complex>
#include <vector>
#include "rade_common.h"
#include "rade_host.h"

typedef std::complex<double> cd;
typedef std::complex<float> cf;

static float2 f2(cf v) { return make_float2(v.real(), v.imag()); }
// exp(j*ang) for a float32 angle, rounded to complex64 (torch.exp / np.exp on a complex64 argument)
static cf cexp32(float ang) { return cf((float)std::cos((double)ang), (float)std::sin((double)ang)); }

static cf pa_limit(cf x) {            // tanh(|x|)*exp(j*angle(x)), float32 (radae/dsp.py:376-377)
  float mag = std::hypot(x.real(), x.imag());
  float ang = std::atan2(x.imag(), x.real());
  float t = std::tanh(mag);
  cf e = cexp32(ang);
  return cf(t * e.real(), t * e.imag());
}

void dsp_tables_host(DspTablesHost &T) {
  const int M = RADE_M, NC = RADE_NC;
  static const float barker13[13] = {1, 1, 1, 1, 1, -1, -1, 1, 1, -1, 1, -1, 1};
  const float two_pi_f = (float)(2.0 * M_PI);
  T.w.resize(NC);
  for (int c = 0; c < NC; c++) T.w[c] = (two_pi_f * (float)(15 + c)) / (float)M;           // radae.py:174 (float32 tensor)
  T.Winv.resize(NC * M); T.Wfwd.resize(M * NC);
  for (int c = 0; c < NC; c++)
    for (int n = 0; n < M; n++) {
      float ang = (float)n * T.w[c];                                                       // float32 product (radae.py:178-179)
      cf e = cexp32(ang);
      T.Winv[c * M + n] = cf(e.real() / (float)M, e.imag() / (float)M);
      T.Wfwd[n * NC + c] = std::conj(e);
    }
  T.P.resize(NC); T.Pend.resize(NC);
  const float sqrt2f = (float)std::sqrt(2.0);
  for (int c = 0; c < NC; c++) {
    T.P[c] = cf(sqrt2f * barker13[c % 13], 0.f);                                           // radae.py:48-56, :182
    T.Pend[c] = (c & 1) ? -T.P[c] : T.P[c];                                                // radae.py:184-185
  }
  T.p.assign(M, cf(0, 0)); T.pend.assign(M, cf(0, 0));
  for (int n = 0; n < M; n++) {
    cf a(0, 0), b(0, 0);
    for (int c = 0; c < NC; c++) { a += T.P[c] * T.Winv[c * M + n]; b += T.Pend[c] * T.Winv[c * M + n]; }
    T.p[n] = a; T.pend[n] = b;                                                             // radae.py:183, :186
  }
  T.pilot_gain = std::pow(10.0, -2.0 / 20.0) * M / std::sqrt((double)NC);                  // radae.py:195-199
  // coarse frequency grid and shifted pilots (radae/dsp.py:163-173), complex128 product stored as csingle
  T.fcoarse.resize(RADE_NFCOARSE); T.p_w.resize(M * RADE_NFCOARSE);
  for (int i = 0; i < RADE_NFCOARSE; i++) {
    double f = -50.0 + 2.5 * i;
    T.fcoarse[i] = (float)f;
    double w = 2 * M_PI * f / RADE_FS;
    for (int n = 0; n < M; n++) {
      cd v = std::exp(cd(0, w * n)) * cd(T.p[n].real(), T.p[n].imag());
      T.p_w[n * RADE_NFCOARSE + i] = cf((float)v.real(), (float)v.imag());
    }
  }
  // symmetric coarse grid: f = +-2.5k Hz, k = 0..20 -> one (cos, sin) pair serves two grid points (ofdm_rx.cu corr6)
  T.cs_tab.assign(M * RADE_CSK, cf(0, 0));
  for (int n = 0; n < M; n++)
    for (int k = 0; k <= 20; k++) {
      double a = 2 * M_PI * (2.5 * k) / RADE_FS * n;
      T.cs_tab[n * RADE_CSK + k] = cf((float)std::cos(a), (float)std::sin(a));
    }
  // Low-rank basis of the coarse grid (AcqTables in rade_common.h): one-sided (Hestenes) Jacobi SVD, in double, of
  //   C[k][m] = cos(w_k (m + 1/2)), S[k][m] = sin(w_k (m + 1/2)),  k = 0..20, m = 0..79, w_k = 2 pi 2.5 k / Fs.
  // Rows are rotated pairwise until orthogonal; then row i = sigma_i v_i^T and C = U^T rows.  The RADE_SRANK strongest rows are
  // the basis, U sigma the expansion coefficients.  The one-sided form keeps the small singular values accurate (the sixth is
  // 6e-8 of the first; the eigenvalues of C C^T would lose it).
  {
    const int NK = 21, NM = RADE_M / 2, R = RADE_SRANK;
    T.srch_basis.assign(NM * 16, 0.f); T.srch_expand.assign(NK * 12, 0.f); T.srch_residual = 0.0;
    for (int fam = 0; fam < 2; fam++) {
      std::vector<double> A(NK * NM), A0, U(NK * NK, 0.0);
      for (int k = 0; k < NK; k++)
        for (int m = 0; m < NM; m++) {
          const double a = 2 * M_PI * (2.5 * k) / RADE_FS * (m + 0.5);
          A[k * NM + m] = fam ? std::sin(a) : std::cos(a);
        }
      A0 = A;
      for (int k = 0; k < NK; k++) U[k * NK + k] = 1.0;
      for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0;
        for (int i = 0; i < NK; i++)
          for (int j = i + 1; j < NK; j++) {
            double al = 0, be = 0, ga = 0;
            for (int m = 0; m < NM; m++) { al += A[i * NM + m] * A[i * NM + m]; be += A[j * NM + m] * A[j * NM + m]; ga += A[i * NM + m] * A[j * NM + m]; }
            if (std::fabs(ga) <= 1e-300 || std::fabs(ga) <= 1e-17 * std::sqrt(al * be)) continue;
            off = std::max(off, std::fabs(ga) / std::sqrt(al * be + 1e-300));
            const double zeta = (be - al) / (2 * ga);
            const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
            const double c = 1 / std::sqrt(1 + t * t), sn = c * t;
            for (int m = 0; m < NM; m++) { const double x = A[i * NM + m], y = A[j * NM + m]; A[i * NM + m] = c * x - sn * y; A[j * NM + m] = sn * x + c * y; }
            for (int m = 0; m < NK; m++) { const double x = U[i * NK + m], y = U[j * NK + m]; U[i * NK + m] = c * x - sn * y; U[j * NK + m] = sn * x + c * y; }
          }
        if (off < 1e-15) break;
      }
      // A0 = U^T A: A0[k][m] = sum_i U[i][k] A[i][m].  Strongest rows first.
      std::vector<std::pair<double, int>> sv;
      for (int i = 0; i < NK; i++) { double n2 = 0; for (int m = 0; m < NM; m++) n2 += A[i * NM + m] * A[i * NM + m]; sv.push_back({std::sqrt(n2), i}); }
      std::sort(sv.begin(), sv.end(), [](const std::pair<double, int> &a, const std::pair<double, int> &b) { return a.first > b.first; });
      std::vector<double> B(NM * R), E(NK * R);
      for (int r = 0; r < R; r++) {
        const int i = sv[r].second; const double sg = sv[r].first;
        const double scale = std::sqrt(sg);                      // split sigma evenly between basis and coefficients
        for (int m = 0; m < NM; m++) B[m * R + r] = A[i * NM + m] / sg * scale;
        for (int k = 0; k < NK; k++) E[k * R + r] = U[i * NK + k] * sg / scale;
      }
      // what the kernels will use is the float32 rounding of both tables: measure the residual of exactly that
      for (int k = 0; k < NK; k++)
        for (int m = 0; m < NM; m++) {
          double acc = 0;
          for (int r = 0; r < R; r++) acc += (double)(float)E[k * R + r] * (double)(float)B[m * R + r];
          T.srch_residual = std::max(T.srch_residual, std::fabs(acc - A0[k * NM + m]));
        }
      for (int m = 0; m < NM; m++)
        for (int r = 0; r < R; r++) T.srch_basis[m * 16 + fam * 8 + r] = (float)B[m * R + r];
      for (int k = 0; k < NK; k++)
        for (int r = 0; r < R; r++) T.srch_expand[k * 12 + fam * 6 + r] = (float)E[k * R + r];
    }
  }
  // LS projectors: Pmat[c] = inv(A^T A) A^T, A = [[1, e^{-j w_{m-1} a}], [1, e^{-j w_m a}], [1, e^{-j w_{m+1} a}]]
  const double a = 0.0025 * RADE_FS;
  T.Pmat.resize(NC * 6); T.eq_rot.resize(NC);
  for (int c = 0; c < NC; c++) {
    int m = c < 1 ? 1 : (c > NC - 2 ? NC - 2 : c);
    cd A[3][2];
    for (int k = 0; k < 3; k++) { A[k][0] = 1.0; A[k][1] = std::exp(cd(0, -(double)T.w[m - 1 + k] * a)); }
    cd G[2][2] = {{0, 0}, {0, 0}};
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 3; k++) G[i][j] += A[k][i] * A[k][j];
    cd det = G[0][0] * G[1][1] - G[0][1] * G[1][0];
    cd Gi[2][2] = {{G[1][1] / det, -G[0][1] / det}, {-G[1][0] / det, G[0][0] / det}};
    for (int i = 0; i < 2; i++) for (int k = 0; k < 3; k++) {
      cd v = Gi[i][0] * A[k][0] + Gi[i][1] * A[k][1];
      T.Pmat[c * 6 + i * 3 + k] = cf((float)v.real(), (float)v.imag());
    }
    T.eq_rot[c] = cexp32(-(T.w[c] * (float)a));                                            // dsp.py:433 (complex64)
  }
  // band-pass filter: float32 scalar arithmetic, left to right (radae_rxe.py:106-107, dsp.py:42-49)
  const float w0 = T.w[0], w29 = T.w[NC - 1];
  const float bw = ((1.2f * (w29 - w0)) * (float)RADE_FS) / two_pi_f;
  const float centre = (((w29 + w0) * (float)RADE_FS) / two_pi_f) / 2.f;
  const float B = bw / (float)RADE_FS;
  const float alpha = (two_pi_f * centre) / (float)RADE_FS;
  T.bpf_bw = bw; T.bpf_centre = centre; T.bpf_alpha = alpha;
  T.bpf_h.resize(RADE_BPF_NTAP);
  for (int i = 0; i < RADE_BPF_NTAP; i++) {
    float n = (float)(i - (RADE_BPF_NTAP - 1) / 2);
    float x = n * B;
    float y = (float)M_PI * (x == 0.f ? 1.0e-20f : x);
    float s = (float)std::sin((double)y) / y;                                              // np.sinc on float32
    T.bpf_h[i] = B * s;
  }
  T.bpf_exp.resize(RADE_NEOO);                      // the receive filter reads [0, 1120), the TX filter up to the 1152-sample EOO frame
  for (int i = 0; i < RADE_NEOO; i++) {
    float arg = (float)(-((double)alpha * (double)(i + 1)));                               // argument rounded to float32 (dsp.py:61)
    T.bpf_exp[i] = cexp32(arg);
  }
  // EOO frame skeleton P E 0 0 0 E (radae.py:208-219)
  T.eoo_base.assign(RADE_NEOO, cf(0, 0));
  const float pg = (float)T.pilot_gain;
  for (int i = 0; i < RADE_SYM; i++) {
    int src = (i < RADE_NCP) ? (M - RADE_NCP + i) : (i - RADE_NCP);
    T.eoo_base[i] = T.p[src];
    T.eoo_base[RADE_SYM + i] = T.pend[src];
    T.eoo_base[RADE_NMF + i] = T.pend[src];
  }
  for (auto &v : T.eoo_base) v = pa_limit(cf(v.real() * pg, v.imag() * pg));
}

int dsp_tables_upload(const DspTablesHost &T, DspTables *D, std::vector<void *> &allocs) {
  auto up = [&](const void *src, size_t bytes) -> void * {
    void *d = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    allocs.push_back(d);
    return d;
  };
#define UPC(field, vec) if (!(D->field = (const float2 *)up(vec.data(), vec.size() * sizeof(cf)))) return -1;
  UPC(Winv, T.Winv) UPC(Wfwd, T.Wfwd) UPC(P, T.P) UPC(Pend, T.Pend) UPC(p, T.p) UPC(pend, T.pend)
  UPC(p_w, T.p_w) UPC(cs_tab, T.cs_tab) UPC(Pmat, T.Pmat) UPC(eq_rot, T.eq_rot) UPC(bpf_exp, T.bpf_exp) UPC(eoo_base, T.eoo_base)
#undef UPC
  if (!(D->bpf_h = (const float *)up(T.bpf_h.data(), T.bpf_h.size() * 4))) return -1;
  if (!(D->fcoarse = (const float *)up(T.fcoarse.data(), T.fcoarse.size() * 4))) return -1;
  {
    std::vector<unsigned char> blob(sizeof(AcqTables), 0);
    AcqTables *A = reinterpret_cast<AcqTables *>(blob.data());
    for (int n = 0; n < RADE_M; n++) {
      const float px = T.p[n].real(), py = T.p[n].imag();
      A->ps4[n] = make_float4(px, py, py, -px);
      A->pcd[n] = make_double2((double)px, -(double)py);
      A->pend[n] = f2(T.pend[n]);
    }
    for (int n = 0; n < RADE_M; n++) {
      const double sc = (n - 79.5) / 80.0;
      double v = 1.0;
      for (int k = 0; k < 10; k++) { A->bk[n][k] = k < 9 ? v : 0.0; v = v * sc / (k + 1); }
    }
    for (int pos = 0; pos < 2; pos++)
      for (int i = 0; i < 24; i++) {
        const double a = 2 * M_PI * (-1.0 + 0.1 * i) / RADE_FS * (79.5 + RADE_NMF * pos);
        A->phd[pos][i] = make_double2(std::cos(a), -std::sin(a));
      }
    for (int pos = 0; pos < 2; pos++)
      for (int i = 0; i < 80; i++) {
        const double a = 2 * M_PI * (-10.0 + 0.25 * i) / RADE_FS * (79.5 + RADE_NMF * pos);
        A->phd10[pos][i] = make_double2(std::cos(a), -std::sin(a));
      }
    memcpy(A->basis, T.srch_basis.data(), sizeof(A->basis));
    memcpy(A->expand, T.srch_expand.data(), sizeof(A->expand));
    if (T.srch_residual > 2e-7) { fprintf(stderr, "libradae_b200: coarse-grid basis residual %.3g\n", T.srch_residual); return -1; }
    if (!(D->acq_tab = (const unsigned char *)up(blob.data(), blob.size()))) return -1;
  }
  D->pilot_gain = (float)T.pilot_gain;
  D->p0_abs = std::abs(T.P[0]);
  return 0;
}

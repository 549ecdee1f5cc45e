// Host-side unit test of the FFT building blocks in tike_b200/csrc/fft.cuh.
// Compiled with nvcc and run on the CPU (no GPU needed): checks the radix
// butterflies, the digit-reversed ordering and the inverse against a naive
// O(N^2) DFT in double precision.
#include <cmath>
#include <complex>
#include <cstdio>
#include <vector>

#include "../../tike_b200/csrc/fft.cuh"
#include "../../tike_b200/csrc/dft32.cuh"

using namespace tb;
typedef std::complex<double> cd;

template <int R>
double test_radix() {
  float2 x[R];
  std::vector<cd> in(R);
  for (int i = 0; i < R; ++i) {
    in[i] = cd(std::sin(1.0 + 3.7 * i), std::cos(0.3 + 1.9 * i * i));
    x[i] = make_float2((float)in[i].real(), (float)in[i].imag());
  }
  dft<R>(x);
  double err = 0;
  for (int k = 0; k < R; ++k) {
    cd acc = 0;
    for (int n = 0; n < R; ++n)
      acc += in[n] * std::polar(1.0, -2.0 * M_PI * n * k / R);
    err = std::max(err, std::abs(acc - cd(x[k].x, x[k].y)));
  }
  return err;
}

// native inverse butterflies (conjugated constants) against the naive inverse DFT
template <int R>
double test_radix_inverse() {
  float2 x[R];
  std::vector<cd> in(R);
  for (int i = 0; i < R; ++i) {
    in[i] = cd(std::cos(0.7 + 2.3 * i), std::sin(1.1 + 0.9 * i * i));
    x[i] = make_float2((float)in[i].real(), (float)in[i].imag());
  }
  idft<R>(x);
  double err = 0;
  for (int k = 0; k < R; ++k) {
    cd acc = 0;
    for (int n = 0; n < R; ++n)
      acc += in[n] * std::polar(1.0, 2.0 * M_PI * n * k / R);
    err = std::max(err, std::abs(acc - cd(x[k].x, x[k].y)));
  }
  return err;
}

template <int N>
double test_2d(double* inv_err) {
  constexpr int P = N + 1;
  std::vector<float2> tile(N * P), tw(N);
  std::vector<cd> in(N * N);
  fill_twiddles<N>(tw.data());
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < N; ++c) {
      in[r * N + c] = cd(std::sin(0.1 * r * c + r), std::cos(0.37 * c - 0.01 * r * r));
      tile[r * P + c] = make_float2((float)in[r * N + c].real(), (float)in[r * N + c].imag());
    }
  fft2_tile<N, false>(tile.data(), tw.data());
  // naive separable DFT
  std::vector<cd> tmp(N * N), out(N * N);
  for (int r = 0; r < N; ++r)
    for (int k = 0; k < N; ++k) {
      cd a = 0;
      for (int c = 0; c < N; ++c) a += in[r * N + c] * std::polar(1.0, -2.0 * M_PI * c * k / N);
      tmp[r * N + k] = a;
    }
  for (int k2 = 0; k2 < N; ++k2)
    for (int k = 0; k < N; ++k) {
      cd a = 0;
      for (int r = 0; r < N; ++r) a += tmp[r * N + k] * std::polar(1.0, -2.0 * M_PI * r * k2 / N);
      out[k2 * N + k] = a;
    }
  double err = 0, nrm = 0;
  for (int ly = 0; ly < N; ++ly)
    for (int lx = 0; lx < N; ++lx) {
      const int ky = loc2freq<N>(ly), kx = loc2freq<N>(lx);
      if (freq2loc<N>(ky) != ly) return 1e9;
      cd got(tile[ly * P + lx].x, tile[ly * P + lx].y);
      err = std::max(err, std::abs(got - out[ky * N + kx]));
      nrm = std::max(nrm, std::abs(out[ky * N + kx]));
    }
  fft2_tile<N, true>(tile.data(), tw.data());
  double e2 = 0;
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < N; ++c) {
      cd got(tile[r * P + c].x / (double)(N * N), tile[r * P + c].y / (double)(N * N));
      e2 = std::max(e2, std::abs(got - in[r * N + c]));
    }
  *inv_err = e2;
  return err / nrm;
}

// radix-32 butterflies of the 512-point transforms (dft32.cuh): slot p of the
// output holds frequency dft32_freq(p); the inverse consumes slot order
double test_radix32(double* inv_err) {
  float2 x[32];
  std::vector<cd> in(32);
  for (int i = 0; i < 32; ++i) {
    in[i] = cd(std::sin(0.4 + 2.1 * i), std::cos(1.3 + 0.7 * i * i));
    x[i] = make_float2((float)in[i].real(), (float)in[i].imag());
  }
  dft32(x);
  double err = 0;
  for (int p = 0; p < 32; ++p) {
    const int k = dft32_freq(p);
    cd acc = 0;
    for (int n = 0; n < 32; ++n) acc += in[n] * std::polar(1.0, -2.0 * M_PI * n * k / 32);
    err = std::max(err, std::abs(acc - cd(x[p].x, x[p].y)));
  }
  idft32(x);
  double e2 = 0;
  for (int n = 0; n < 32; ++n)
    e2 = std::max(e2, std::abs(cd(x[n].x / 32.0, x[n].y / 32.0) - in[n]));
  *inv_err = e2;
  return err;
}

// The three-pass 128 x 128 transform of rpie_p3.cu, pass by pass with the same
// butterflies and twiddles (column plan 8 x 16, row plan 4 x 2 x 16), one
// "thread" at a time: checks the ownerships and the slot -> frequency maps the
// kernel relies on (row slot r holds frequency (r >> 4) + 8 (r & 15), column
// slot 32 a1 + 16 b1 + p1 holds a1 + 4 b1 + 8 p1), and the inverse.
double test_three_pass(double* inv_err) {
  constexpr int N = 128;
  std::vector<float2> t(N * N), tw(N);
  std::vector<cd> in(N * N);
  fill_twiddles<N>(tw.data());
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < N; ++c) {
      in[r * N + c] = cd(std::sin(0.13 * r * c + c), std::cos(0.29 * r - 0.02 * c * c));
      t[r * N + c] = make_float2((float)in[r * N + c].real(), (float)in[r * N + c].imag());
    }
  auto at = [&](int r, int c) -> float2& { return t[r * N + c]; };
  // pass 1: rows n2 + 16 k, columns m + 32 a
  for (int n2 = 0; n2 < 16; ++n2)
    for (int m = 0; m < 32; ++m) {
      float2 v[4][8];
      for (int a = 0; a < 4; ++a) {
        for (int k = 0; k < 8; ++k) v[a][k] = at(n2 + 16 * k, m + 32 * a);
        dft<8>(v[a]);
        for (int k = 1; k < 8; ++k) v[a][k] = cmul(v[a][k], tw[n2 * k]);
      }
      for (int k = 0; k < 8; ++k) {
        float2 q[4] = {v[0][k], v[1][k], v[2][k], v[3][k]};
        dft<4>(q);
        for (int a = 0; a < 4; ++a) at(n2 + 16 * k, m + 32 * a) = a ? cmul(q[a], tw[(m * a) & 127]) : q[0];
      }
    }
  // pass 2: rows 16 k1 + n2, columns 32 a1 + 16 b + p
  for (int k1 = 0; k1 < 8; ++k1)
    for (int a1 = 0; a1 < 4; ++a1)
      for (int p = 0; p < 16; ++p) {
        float2 u[2][16];
        for (int b = 0; b < 2; ++b) {
          for (int n = 0; n < 16; ++n) u[b][n] = at(16 * k1 + n, 32 * a1 + 16 * b + p);
          dft<16>(u[b]);
        }
        for (int n = 0; n < 16; ++n) {
          at(16 * k1 + n, 32 * a1 + p) = cadd(u[0][n], u[1][n]);
          at(16 * k1 + n, 32 * a1 + 16 + p) = cmul(csub(u[0][n], u[1][n]), tw[4 * p]);
        }
      }
  // pass 3: row r, 16 consecutive columns
  for (int r = 0; r < N; ++r)
    for (int B = 0; B < 8; ++B) {
      float2 z[16];
      for (int p = 0; p < 16; ++p) z[p] = at(r, 16 * B + p);
      dft<16>(z);
      for (int p = 0; p < 16; ++p) at(r, 16 * B + p) = z[p];
    }
  // naive separable DFT
  std::vector<cd> tmp(N * N), out(N * N);
  for (int r = 0; r < N; ++r)
    for (int k = 0; k < N; ++k) {
      cd a = 0;
      for (int c = 0; c < N; ++c) a += in[r * N + c] * std::polar(1.0, -2.0 * M_PI * c * k / N);
      tmp[r * N + k] = a;
    }
  for (int k2 = 0; k2 < N; ++k2)
    for (int k = 0; k < N; ++k) {
      cd a = 0;
      for (int r = 0; r < N; ++r) a += tmp[r * N + k] * std::polar(1.0, -2.0 * M_PI * r * k2 / N);
      out[k2 * N + k] = a;
    }
  double err = 0, nrm = 0;
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < N; ++c) {
      const int fr = (r >> 4) + 8 * (r & 15);
      const int fc = (c >> 5) + 4 * ((c >> 4) & 1) + 8 * (c & 15);
      err = std::max(err, std::abs(cd(at(r, c).x, at(r, c).y) - out[fr * N + fc]));
      nrm = std::max(nrm, std::abs(out[fr * N + fc]));
    }
  // inverse: the mirror image
  for (int r = 0; r < N; ++r)
    for (int B = 0; B < 8; ++B) {
      float2 z[16];
      for (int p = 0; p < 16; ++p) z[p] = at(r, 16 * B + p);
      idft<16>(z);
      for (int p = 0; p < 16; ++p) at(r, 16 * B + p) = z[p];
    }
  for (int k1 = 0; k1 < 8; ++k1)
    for (int a1 = 0; a1 < 4; ++a1)
      for (int p = 0; p < 16; ++p) {
        float2 u[2][16];
        for (int n = 0; n < 16; ++n) {
          const float2 x0 = at(16 * k1 + n, 32 * a1 + p);
          const float2 x1 = cmulc(tw[4 * p], at(16 * k1 + n, 32 * a1 + 16 + p));
          u[0][n] = cadd(x0, x1);
          u[1][n] = csub(x0, x1);
        }
        idft<16>(u[0]);
        idft<16>(u[1]);
        for (int b = 0; b < 2; ++b)
          for (int n = 0; n < 16; ++n) at(16 * k1 + n, 32 * a1 + 16 * b + p) = u[b][n];
      }
  for (int n2 = 0; n2 < 16; ++n2)
    for (int m = 0; m < 32; ++m) {
      float2 v[4][8];
      for (int k = 0; k < 8; ++k) {
        float2 q[4];
        for (int a = 0; a < 4; ++a)
          q[a] = a ? cmulc(tw[(m * a) & 127], at(n2 + 16 * k, m + 32 * a)) : at(n2 + 16 * k, m);
        idft<4>(q);
        for (int a = 0; a < 4; ++a) v[a][k] = q[a];
      }
      for (int a = 0; a < 4; ++a) {
        for (int k = 1; k < 8; ++k) v[a][k] = cmulc(tw[n2 * k], v[a][k]);
        idft<8>(v[a]);
        for (int k = 0; k < 8; ++k) at(n2 + 16 * k, m + 32 * a) = v[a][k];
      }
    }
  double e2 = 0;
  for (int i = 0; i < N * N; ++i)
    e2 = std::max(e2, std::abs(cd(t[i].x / (double)(N * N), t[i].y / (double)(N * N)) - in[i]));
  *inv_err = e2;
  return err / nrm;
}

// The two-stage 1-D plans of the register-resident large-detector kernels
// (large_k2r.cu, large_k13r.cu): N = 16 x R2 with R2 = 16 (N = 256) or 32 (N =
// 512, dft32).  Stage 1: radix-16 over elements n2 + R2 k, twiddle w_N^(n2 k1);
// stage 2: radix-R2 over elements R2 k1 + n.  Slot R2 k1 + p holds frequency
// k1 + 16 f(p), f(p) = p (R2 = 16) or dft32_freq(p) (R2 = 32) -- the map K2 uses
// to find the measured pixel of a far-field value.
template <int N>
double test_two_stage(double* inv_err) {
  constexpr int R2 = N / 16;
  std::vector<float2> t(N), tw(N);
  std::vector<cd> in(N);
  fill_twiddles<N>(tw.data());
  for (int i = 0; i < N; ++i) {
    in[i] = cd(std::sin(0.31 * i + 0.002 * i * i), std::cos(1.7 * i));
    t[i] = make_float2((float)in[i].real(), (float)in[i].imag());
  }
  for (int n2 = 0; n2 < R2; ++n2) {
    float2 x[16];
    for (int k = 0; k < 16; ++k) x[k] = t[n2 + R2 * k];
    dft<16>(x);
    for (int k = 0; k < 16; ++k) t[n2 + R2 * k] = k ? cmul(x[k], tw[n2 * k]) : x[0];
  }
  for (int k1 = 0; k1 < 16; ++k1) {
    if constexpr (R2 == 16) {
      float2 y[16];
      for (int n = 0; n < 16; ++n) y[n] = t[16 * k1 + n];
      dft<16>(y);
      for (int n = 0; n < 16; ++n) t[16 * k1 + n] = y[n];
    } else {
      float2 y[32];
      for (int n = 0; n < 32; ++n) y[n] = t[32 * k1 + n];
      dft32(y);
      for (int n = 0; n < 32; ++n) t[32 * k1 + n] = y[n];
    }
  }
  double err = 0, nrm = 0;
  for (int s = 0; s < N; ++s) {
    const int k1 = s / R2, p = s % R2;
    const int f = k1 + 16 * (R2 == 16 ? p : dft32_freq(p));
    cd acc = 0;
    for (int n = 0; n < N; ++n) acc += in[n] * std::polar(1.0, -2.0 * M_PI * n * f / N);
    err = std::max(err, std::abs(acc - cd(t[s].x, t[s].y)));
    nrm = std::max(nrm, std::abs(acc));
  }
  // inverse: stage 2 first, then conjugate twiddles and stage 1
  for (int k1 = 0; k1 < 16; ++k1) {
    if constexpr (R2 == 16) {
      float2 y[16];
      for (int n = 0; n < 16; ++n) y[n] = t[16 * k1 + n];
      idft<16>(y);
      for (int n = 0; n < 16; ++n) t[16 * k1 + n] = y[n];
    } else {
      float2 y[32];
      for (int n = 0; n < 32; ++n) y[n] = t[32 * k1 + n];
      idft32(y);
      for (int n = 0; n < 32; ++n) t[32 * k1 + n] = y[n];
    }
  }
  for (int n2 = 0; n2 < R2; ++n2) {
    float2 x[16];
    for (int k = 0; k < 16; ++k) x[k] = k ? cmulc(tw[n2 * k], t[n2 + R2 * k]) : t[n2];
    idft<16>(x);
    for (int k = 0; k < 16; ++k) t[n2 + R2 * k] = x[k];
  }
  double e2 = 0;
  for (int i = 0; i < N; ++i)
    e2 = std::max(e2, std::abs(cd(t[i].x / (double)N, t[i].y / (double)N) - in[i]));
  *inv_err = e2;
  return err / nrm;
}

int main() {
  int fail = 0;
  double e;
  e = test_radix<2>();  printf("radix2  err %.3e\n", e); fail |= e > 1e-5;
  e = test_radix<4>();  printf("radix4  err %.3e\n", e); fail |= e > 1e-5;
  e = test_radix<8>();  printf("radix8  err %.3e\n", e); fail |= e > 1e-5;
  e = test_radix<16>(); printf("radix16 err %.3e\n", e); fail |= e > 1e-5;
  e = test_radix_inverse<2>();  printf("iradix2  err %.3e\n", e); fail |= e > 1e-5;
  e = test_radix_inverse<4>();  printf("iradix4  err %.3e\n", e); fail |= e > 1e-5;
  e = test_radix_inverse<8>();  printf("iradix8  err %.3e\n", e); fail |= e > 1e-5;
  e = test_radix_inverse<16>(); printf("iradix16 err %.3e\n", e); fail |= e > 1e-5;
  double ie;
#define T2(N) e = test_2d<N>(&ie); printf("fft2 %4d rel err %.3e  roundtrip err %.3e\n", N, e, ie); fail |= (e > 2e-6) | (ie > 2e-5);
  T2(16) T2(32) T2(64) T2(128) T2(256)
  e = test_radix32(&ie); printf("radix32 err %.3e  roundtrip err %.3e\n", e, ie); fail |= (e > 2e-5) | (ie > 2e-6);
  e = test_three_pass(&ie); printf("three-pass 128 rel err %.3e  roundtrip err %.3e\n", e, ie); fail |= (e > 2e-6) | (ie > 2e-5);
  e = test_two_stage<256>(&ie); printf("two-stage 256 rel err %.3e  roundtrip err %.3e\n", e, ie); fail |= (e > 2e-6) | (ie > 2e-6);
  e = test_two_stage<512>(&ie); printf("two-stage 512 rel err %.3e  roundtrip err %.3e\n", e, ie); fail |= (e > 2e-6) | (ie > 2e-6);
  // 1-D only for the big plans (2-D naive would be slow): use 1 vector
  printf(fail ? "FAIL\n" : "PASS\n");
  return fail;
}

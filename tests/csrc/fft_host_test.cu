// Host-side unit test of the FFT building blocks in tike_b200/csrc/fft.cuh.
// Compiled with nvcc and run on the CPU (no GPU needed): checks the radix
// butterflies, the digit-reversed ordering and the inverse against a naive
// O(N^2) DFT in double precision.
#include <cmath>
#include <complex>
#include <cstdio>
#include <vector>

#include "../../tike_b200/csrc/fft.cuh"

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
  // 1-D only for the big plans (2-D naive would be slow): use 1 vector
  printf(fail ? "FAIL\n" : "PASS\n");
  return fail;
}

// Cubic-spline baseline of the reference's evaluation (3_test_cubic_interpolation.py:32-58) on the device.
//
// The reference zeroes the masked frames, turns every exact 0 into NaN and calls, per keypoint and coordinate,
// pandas Series.interpolate(method="cubicspline", limit_direction="both"), i.e. scipy.interpolate.CubicSpline (not-a-knot
// end conditions, extrapolation on both sides) through the remaining samples, evaluated at the missing frames
// (~0.27 s per T = 256 sequence on the host).  Here: one thread per (sequence, keypoint, coordinate) series, fp64.
//   n >= 4 knots: first derivatives s from the tridiagonal system scipy builds (interior rows
//                 dx[i] s[i-1] + 2 (dx[i-1] + dx[i]) s[i] + dx[i-1] s[i+1] = 3 (dx[i] m[i-1] + dx[i-1] m[i]), m = slopes,
//                 not-a-knot first / last rows), eliminated top to bottom, then the Hermite cubic of the interval;
//   n == 3: the parabola through the three points;  n == 2: the straight line;  n == 1: that constant (scipy raises);
//   n == 0: zeros (np.nan_to_num).
#include "common.cuh"

namespace kit {

template <int MAXT>
__global__ void __launch_bounds__(128) cubic_fill_kernel(const float* __restrict__ data, const float* __restrict__ mask,
                                                         float* __restrict__ out, int64_t n_series, int T1, int K) {
  pdl_grid_sync();
  const int64_t sidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (sidx >= n_series) return;
  const int64_t b = sidx / (2 * K);
  const int kc = (int)(sidx - b * 2 * K);   // k * 2 + coordinate
  const float* src = data + (b * T1) * (int64_t)(2 * K) + kc;
  float* dst = out + (b * T1) * (int64_t)(2 * K) + kc;
  const float* mrow = mask + b * T1;
  const int64_t stride = 2 * K;
  short xs[MAXT];
  double ys[MAXT], cp[MAXT], dp[MAXT];
  int n = 0;
  for (int t = 0; t < T1; ++t) {
    const float v = src[t * stride];
    if (mrow[t] != 1.f && v != 0.f && v == v) {
      xs[n] = (short)t;
      ys[n] = (double)v;
      ++n;
    }
  }
  if (n <= 1) {
    const float fill = n == 1 ? (float)ys[0] : 0.f;
    for (int t = 0; t < T1; ++t) dst[t * stride] = fill;
    return;
  }
  // first derivatives at the knots -> dp[]
  if (n == 2) {
    dp[0] = dp[1] = (ys[1] - ys[0]) / (double)(xs[1] - xs[0]);
  } else if (n == 3) {
    const double h0 = xs[1] - xs[0], h1 = xs[2] - xs[1];
    const double m0 = (ys[1] - ys[0]) / h0, m1 = (ys[2] - ys[1]) / h1;
    // [1 1 0; h1 2(h0+h1) h0; 0 1 1] s = [2 m0, 3 (h0 m1 + h1 m0), 2 m1]
    const double s1 = (3.0 * (h0 * m1 + h1 * m0) - 2.0 * h1 * m0 - 2.0 * h0 * m1) / (h0 + h1);
    dp[1] = s1;
    dp[0] = 2.0 * m0 - s1;
    dp[2] = 2.0 * m1 - s1;
  } else {
    {   // not-a-knot first row: dx1 s0 + (x2 - x0) s1 = ((dx0 + 2 d) dx1 m0 + dx0^2 m1) / d
      const double h0 = xs[1] - xs[0], h1 = xs[2] - xs[1], d = xs[2] - xs[0];
      const double m0 = (ys[1] - ys[0]) / h0, m1 = (ys[2] - ys[1]) / h1;
      cp[0] = d / h1;
      dp[0] = ((h0 + 2.0 * d) * h1 * m0 + h0 * h0 * m1) / d / h1;
    }
    for (int i = 1; i < n - 1; ++i) {
      const double hl = xs[i] - xs[i - 1], hr = xs[i + 1] - xs[i];
      const double ml = (ys[i] - ys[i - 1]) / hl, mr = (ys[i + 1] - ys[i]) / hr;
      const double a = hr, bb = 2.0 * (hl + hr), c = hl, r = 3.0 * (hr * ml + hl * mr);
      const double piv = bb - a * cp[i - 1];
      cp[i] = c / piv;
      dp[i] = (r - a * dp[i - 1]) / piv;
    }
    {   // not-a-knot last row: (x[-1] - x[-3]) s[-2] + dx[-2] s[-1] = (dx[-1]^2 m[-2] + (2 d + dx[-1]) dx[-2] m[-1]) / d
      const int i = n - 1;
      const double hl = xs[i - 1] - xs[i - 2], hr = xs[i] - xs[i - 1], d = xs[i] - xs[i - 2];
      const double ml = (ys[i - 1] - ys[i - 2]) / hl, mr = (ys[i] - ys[i - 1]) / hr;
      const double a = d, bb = hl, r = (hr * hr * ml + (2.0 * d + hr) * hl * mr) / d;
      const double piv = bb - a * cp[i - 1];
      dp[i] = (r - a * dp[i - 1]) / piv;
    }
    for (int i = n - 2; i >= 0; --i) dp[i] -= cp[i] * dp[i + 1];
  }
  // evaluation: knots keep their samples, every other frame takes the cubic of its interval (the end intervals extrapolate)
  int j = 0;
  for (int t = 0; t < T1; ++t) {
    while (j < n - 2 && t >= xs[j + 1]) ++j;
    float v;
    if (t == xs[j]) v = (float)ys[j];
    else if (t == xs[j + 1]) v = (float)ys[j + 1];
    else {
      const double h = xs[j + 1] - xs[j], m = (ys[j + 1] - ys[j]) / h;
      const double tt = (dp[j] + dp[j + 1] - 2.0 * m) / h;
      const double c3 = tt / h, c2 = (m - dp[j]) / h - tt;
      const double u = (double)(t - xs[j]);
      v = (float)(ys[j] + u * (dp[j] + u * (c2 + u * c3)));
    }
    dst[t * stride] = v;
  }
}

}  // namespace kit

using namespace kit;

extern "C" int kit_cubic_interpolate(const float* data, const float* mask, float* out, int32_t B, int32_t T1, int32_t K, void* stream) {
  KIT_REQUIRE(data && mask && out && B > 0 && T1 > 0 && K > 0, "kit_cubic_interpolate: bad arguments");
  KIT_REQUIRE(T1 <= 1040, "kit_cubic_interpolate: at most 1040 frames per sequence (got %d)", T1);
  const int64_t n_series = (int64_t)B * K * 2;
  const dim3 grid((unsigned)ceil_div(n_series, 128)), block(128);
  cudaStream_t st = (cudaStream_t)stream;
  if (T1 <= 80) launch_kernel(cubic_fill_kernel<80>, grid, block, 0, st, data, mask, out, n_series, (int)T1, (int)K);
  else if (T1 <= 272) launch_kernel(cubic_fill_kernel<272>, grid, block, 0, st, data, mask, out, n_series, (int)T1, (int)K);
  else if (T1 <= 528) launch_kernel(cubic_fill_kernel<528>, grid, block, 0, st, data, mask, out, n_series, (int)T1, (int)K);
  else launch_kernel(cubic_fill_kernel<1040>, grid, block, 0, st, data, mask, out, n_series, (int)T1, (int)K);
  KIT_LAUNCH_CHECK();
  return KIT_OK;
}

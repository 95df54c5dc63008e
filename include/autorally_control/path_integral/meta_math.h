// meta_math.h -- compile-time sizes of an MLP given its layer widths (API of PI/meta_math.h:13-46):
// param_counter(6,32,32,4) = 1412, layer_counter = 4, neuron_counter = 32.
#ifndef META_MATH_
#define META_MATH_
#include <initializer_list>

namespace mppi_meta {
constexpr int params_of(const int *w, int n) {
  int total = n == 1 ? w[0] : 0;
  for (int i = 0; i + 1 < n; i++) total += (w[i] + 1) * w[i + 1];
  return total;
}
constexpr int widest_of(const int *w, int n) {
  int m = w[0];
  for (int i = 1; i < n; i++) m = w[i] > m ? w[i] : m;
  return m;
}
}  // namespace mppi_meta

template <typename... Args>
constexpr int param_counter(int first, Args... args) {
  const int w[] = {first, args...};
  return mppi_meta::params_of(w, 1 + (int)sizeof...(Args));
}

template <typename... Args>
constexpr int layer_counter(int, Args... args) {
  return 1 + (int)sizeof...(args);
}

template <typename... Args>
constexpr int neuron_counter(int first, Args... args) {
  const int w[] = {first, args...};
  return mppi_meta::widest_of(w, 1 + (int)sizeof...(Args));
}

#endif

// stdrand.cpp -- TEST INFRASTRUCTURE (oracle/): exposes libstdc++'s std::default_random_engine
// streams so the Python tests can regenerate the reference tests' random inputs exactly:
//   * x0 ~ U(0,1) in DoF order      tests/hierarchy_driver.cc:153-164, tests/test_hierarchy_device.cu:293-297
//   * x_ref ~ N(10,2)               tests/test_direct_solver_device.cu:49-53
//   * 5 column draws per row, seed = row   tests/test_sparse_matrix_device.cu:45-57
#include <cstdint>
#include <random>

extern "C"
{
  void stdrand_uniform01(unsigned seed, int64_t n, double *out)
  {
    std::default_random_engine generator(seed);
    std::uniform_real_distribution<double> distribution(0., 1.);
    for (int64_t i = 0; i < n; ++i)
      out[i] = distribution(generator);
  }

  // draws only for entries with skip[i] == 0 (the driver does not draw for constrained DoFs)
  void stdrand_uniform01_masked(unsigned seed, int64_t n, const unsigned char *skip, double *out)
  {
    std::default_random_engine generator(seed);
    std::uniform_real_distribution<double> distribution(0., 1.);
    for (int64_t i = 0; i < n; ++i)
      out[i] = skip[i] ? 0. : distribution(generator);
  }

  void stdrand_normal(unsigned seed, double mean, double stddev, int64_t n, double *out)
  {
    std::default_random_engine generator(seed);
    std::normal_distribution<> distribution(mean, stddev);
    for (int64_t i = 0; i < n; ++i)
      out[i] = distribution(generator);
  }

  void stdrand_uniform_int(unsigned seed, int lo, int hi, int64_t n, int *out)
  {
    std::default_random_engine generator(seed);
    std::uniform_int_distribution<int> distribution(lo, hi);
    for (int64_t i = 0; i < n; ++i)
      out[i] = distribution(generator);
  }
}

// oracle/eigen_shim/deterministic_random_device.h — TEST INFRASTRUCTURE.
//
// Force-included (g++ -include) in front of the UNMODIFIED /root/reference/lstm.cc when it is built as
// oracle/_ref/lstm_ref.  The reference seeds every generator from std::random_device (R/lstm.cc:370-371, 309-310),
// which makes it irreproducible; this header renames `random_device` to a counter so that the k-th device the
// program constructs returns REF_SEED + k (REF_SEED from the environment, default 0) — the same convention the
// oracle and the C ABI use (lstm_init_params: k-th tensor seeded seed + k).
#ifndef ORACLE_DETERMINISTIC_RANDOM_DEVICE_H_
#define ORACLE_DETERMINISTIC_RANDOM_DEVICE_H_
#include <cstdlib>
#include <random>

namespace std {
class shim_random_device {
 public:
  typedef unsigned int result_type;
  shim_random_device() : value_(base() + counter()++) {}
  result_type operator()() { return value_; }

 private:
  static unsigned int base() {
    const char* s = getenv("REF_SEED");
    return s ? (unsigned int)strtoul(s, 0, 10) : 0u;
  }
  static unsigned int& counter() { static unsigned int k = 0; return k; }
  unsigned int value_;
};
}  // namespace std
#define random_device shim_random_device
#endif

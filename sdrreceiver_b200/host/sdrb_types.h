// Shared by the facade headers: the sample type the reference uses everywhere and the
// exception the facades throw when a C-ABI call fails (the reference has no error returns;
// a failed CUDA call here must not pass silently -- there is no CPU path to fall back to).
#ifndef SDRB_HOST_TYPES_H
#define SDRB_HOST_TYPES_H
#include <complex>
#include <stdexcept>
#include <string>
#include <vector>

typedef std::complex<float> cpx_typef;

namespace sdrb_host {
struct Error : public std::runtime_error {
    explicit Error(const std::string &what) : std::runtime_error(what) {}
};
void check(int rc, const char *what);          // throws Error with sdrb_last_error() when rc != 0
void *dev_alloc(size_t bytes);                 // cudaMalloc / cudaFree / cudaMemcpy wrappers so that
void dev_free(void *p);                        // code including these headers needs no CUDA headers
void to_dev(void *dst, const void *src, size_t bytes);
void to_host(void *dst, const void *src, size_t bytes);
void dev_zero(void *dst, size_t bytes);
}  // namespace sdrb_host
#endif

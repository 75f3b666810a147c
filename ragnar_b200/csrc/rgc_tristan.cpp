// placeholder — replaced by the HDF5 streaming reader
#include "rgc_internal.hpp"
using namespace rgc;
extern "C" {
int rgc_tristan_read_particles(const char*, size_t, unsigned, size_t, size_t, size_t, int, int,
                               rgc_particles_t**, size_t*, size_t*) {
  return fail(RGC_ERR_IO, "Tristan-v2 reader not built yet");
}
int rgc_tristan_write_species(const char*, size_t, unsigned, size_t, int, const float* const*,
                              int) {
  return fail(RGC_ERR_IO, "Tristan-v2 writer not built yet");
}
}

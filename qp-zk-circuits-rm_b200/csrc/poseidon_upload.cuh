// Host side of the Poseidon tables: derives them (poseidon_tables.hpp) and uploads them to the
// __constant__ / __device__ symbols the kernels read (poseidon.cuh). One call per device.
#pragma once
#include <cuda_runtime.h>

#include "poseidon.cuh"
#include "poseidon_tables.hpp"

namespace qpzk {

static inline cudaError_t poseidon_upload_tables(const PoseidonTablesHost& T) {
#define QPZK_UP(sym, src, bytes)                                  \
  do {                                                            \
    cudaError_t e_ = cudaMemcpyToSymbol(sym, src, bytes);         \
    if (e_ != cudaSuccess) return e_;                             \
  } while (0)
  QPZK_UP(c_rc, T.rc, sizeof T.rc);
  {
    static u64 rc_padded[372];
    for (int i = 0; i < 372; i++) rc_padded[i] = i < 360 ? T.rc[i] : 0;
    QPZK_UP(g_rc, rc_padded, sizeof rc_padded);
  }
#if PV_COOP_LINEAR
  QPZK_UP(g_lin_p, T.lin_p, sizeof T.lin_p);
  QPZK_UP(g_lin_c, T.lin_c, sizeof T.lin_c);
  QPZK_UP(g_lin_coef, T.lin_coef, sizeof T.lin_coef);
#endif
  QPZK_UP(c_fast_first, T.fast_first, sizeof T.fast_first);
  QPZK_UP(c_fast_rc, T.fast_rc, sizeof T.fast_rc);
  QPZK_UP(c_fast_init, T.fast_init, sizeof T.fast_init);
  QPZK_UP(c_fast_w_hat, T.fast_w_hat, sizeof T.fast_w_hat);
  QPZK_UP(c_fast_v, T.fast_v, sizeof T.fast_v);
  QPZK_UP(c_h_rc, T.h_rc, sizeof T.h_rc);
  QPZK_UP(c_h_init, T.h_init, sizeof T.h_init);
  QPZK_UP(c_h_w_hat, T.h_w_hat, sizeof T.h_w_hat);
  QPZK_UP(c_h_v, T.h_v, sizeof T.h_v);
  u32 circ[12];
  for (int i = 0; i < 12; i++) circ[i] = (u32)kMdsCirc[i];
  u32 diag0 = (u32)kMdsDiag0;
  QPZK_UP(c_mds_circ, circ, sizeof circ);
  QPZK_UP(c_mds_diag0, &diag0, sizeof diag0);
#if PV_MDS_F64
  double circ_d[12];
  for (int i = 0; i < 12; i++) circ_d[i] = (double)kMdsCirc[i];
  QPZK_UP(c_mds_circ_d, circ_d, sizeof circ_d);
  static double next_rc[QPZK_MDS_LAYERS_MAX][2][12];
  poseidon_next_rc_f64(T, next_rc, PV_MDS_SPLIT != 0);
  double half_d[12];
  poseidon_mds_half_f64(half_d);
  QPZK_UP(c_mds_half_d, half_d, sizeof half_d);
  QPZK_UP(c_mds_next_rc_d, next_rc, sizeof next_rc);
#endif
#undef QPZK_UP
  return cudaSuccess;
}

}  // namespace qpzk

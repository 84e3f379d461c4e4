"""spmv_acc_b200 — B200-native fp64 CSR SpMV engine (y = alpha*A*x + beta*y), the `cuda-b200` kernel strategy for
hpcde/spmv-acc.

The hot path is hand-written sm_100a CUDA behind the C ABI in ``include/spmv_b200.h`` (``lib/libspmv_b200.so``);
this package is the host-side mirror of the reference's operator interface plus measurement helpers. Importing the
package does not load the CUDA library; the first call does, and fails loudly if it has not been built.
"""
from .api import (CsrDesc, HostMatrix, PeerBuffer, SpmvB200Error, SpmvPlan, cache_invalidate, cache_size, col_block_bitmap, coo_to_csr,  # noqa: F401
                  host_spmv, make_options, HaloLoop, cache_revalidations, enable_peer_access, operation_none, operation_transpose, shard_bounds, sparse_csr_spmv,
                  sparse_spmv, FLAG_BETA0_SKIP_Y, FLAG_DIRECT, FLAG_L2_PERSIST_X, FLAG_NO_DIRECT, FLAG_NO_TMA, FLAG_NO_XSTAGE)

__version__ = "0.1.0"

"""Host-side SETUP path (problem generation, agglomeration, local eigensolves, restriction
assembly, Galerkin product).  Per the north star this stays on the host and only produces the CSR
operators that are uploaded once; nothing here is on the V-cycle hot path and nothing here is a
fallback for it."""
from .amge import (aggregate_restrictor, block_agglomerates, build_multilevel, build_restrictor, galerkin, lowest_eigenpairs, restriction_from_local,  # noqa: F401
                   transpose)  # noqa: F401
from .problems import (HostCSR, LaplaceProblem, assemble, boundary_mask, csr_from_dealii_sparse_matrix,  # noqa: F401
                       coefficient_table, material_value, num_threads, reference_matrices,
                       set_num_threads)
from .partition import LocalPart, coarse_dd_plan, make_parts, partition_two_level, slab_row_ranges  # noqa: F401,E402
from .slab import build_slab_part  # noqa: F401,E402
from . import mmio  # noqa: F401,E402

"""monocle3_b200 -- B200-native preprocess_cds PCA path (drop-in for monocle3's
estimate_size_factors / normalized_counts / preprocess_cds / preprocess_transform).

Importing the package does not load the CUDA library; the first call that needs
the device does (and raises if libm3b.so is not built or no B200 is visible --
there is no CPU fallback)."""
from .api import (CellDataSet, aggregate_gene_expression, align_beta_transform, align_cds, load_transform_models,  # noqa: F401
                  model_matrix, save_transform_models)
from .engine import CSCBlocks  # noqa: F401
from .api import (CellDataSet, aggregate_gene_expression, cluster_cells_make_graph, detect_genes, estimate_size_factors,  # noqa: F401
                  get_context, jaccard_coeff, load_a549, new_cell_data_set, normalized_counts, preprocess_cds, preprocess_transform,
                  set_context, set_size_factors, size_factors)

__all__ = ["align_cds", "align_beta_transform", "save_transform_models", "load_transform_models", "model_matrix",
           "CSCBlocks", "CellDataSet", "aggregate_gene_expression", "cluster_cells_make_graph", "detect_genes", "estimate_size_factors", "get_context", "jaccard_coeff", "load_a549", "new_cell_data_set",
           "normalized_counts", "preprocess_cds", "preprocess_transform", "set_context", "set_size_factors",
           "size_factors"]

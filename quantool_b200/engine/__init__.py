"""Host-side orchestration of the sm_100a kernels (one module per method)."""

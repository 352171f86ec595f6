"""TEST INFRASTRUCTURE ONLY: CPU oracle for the EOFluxVAE hot path (see eovae_oracle.py header)."""

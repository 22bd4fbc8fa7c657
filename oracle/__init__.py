"""TEST INFRASTRUCTURE ONLY: CPU oracle for the SOC photon-packet hot path (see soc_oracle.h)."""

"""TEST INFRASTRUCTURE ONLY: CPU oracle for the PPNet EDaGe-PP hot path (see ppnet_oracle.py)."""

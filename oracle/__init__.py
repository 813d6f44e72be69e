"""CPU oracle for the DRR hot path -- TEST INFRASTRUCTURE, not product code (see oracle/drr_oracle.c)."""

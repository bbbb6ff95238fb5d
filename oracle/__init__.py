"""CPU oracle for the kdcc hot path -- test infrastructure only (see kdcc_oracle.c)."""

"""CPU oracle (test infrastructure only — see oracle/pps_oracle.py)."""

"""Tie proofs for the ranking parity tests: the helpers live in oracle/parity.py (test infrastructure shared with
bench.py's oracle-subset checks)."""
from oracle.parity import (TIE_RTOL, assert_counts_exact_on_own_distances, assert_first_rank_parity,  # noqa: F401
                           assert_topk_parity, subset_check)

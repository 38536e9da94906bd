"""CPU oracle — test infrastructure only (see omb_oracle.cpp header)."""

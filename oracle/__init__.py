"""CPU oracle for the hot path.  TEST INFRASTRUCTURE ONLY -- see oracle/np_oracle.py."""

"""CPU oracle (test infrastructure): see oracle/cgpe_oracle.c.  PARITY UNPINNED against ESPResSo."""

"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/b32_oracle.h)."""

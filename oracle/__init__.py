"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/wl_oracle.cpp).  Never imported by waterlily_b200."""

"""CPU oracle — TEST INFRASTRUCTURE ONLY (see phi3_oracle.py header). Not part of the product path."""

"""CPU oracle — TEST INFRASTRUCTURE ONLY (see oracle/dart_oracle.c header).

Nothing under dart_env_b200/ may import this package."""

"""TEST INFRASTRUCTURE ONLY -- CPU checkers; see oracle/README.md."""

"""CPU oracle (test infrastructure only; see ppo_oracle.c header)."""

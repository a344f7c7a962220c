"""Job launcher: ``deepspeed`` CLI (runner) -> per-node ``launch`` -> one process per GPU."""

"""CPU oracle of the cUDE loss path — TEST INFRASTRUCTURE ONLY (see cude_oracle.cpp header).
Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm."""

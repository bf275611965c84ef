"""Oracle package — CPU restatements of the reference used ONLY by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs. Never imported by the product path."""

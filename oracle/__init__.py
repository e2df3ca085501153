"""TEST INFRASTRUCTURE: the CPU oracle.  Importable only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never from misaki_render_b200."""

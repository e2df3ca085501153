"""Synthetic inputs for the tests, bench.py and smoke(): the BASELINE.json configurations.

The reference ships no mesh assets (its .gitignore excludes assets/*/*), so the Cornell box is
re-authored from the public Cornell data and the bunny/teapot/10M-triangle meshes are seeded
procedural stand-ins of the stated size.
"""

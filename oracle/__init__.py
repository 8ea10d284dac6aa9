"""CPU oracle for the TrackDLO registration path -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; the product package (trackdlo_b200) never imports it.  PINNED against the reference's own compiled
sources (oracle/_ref, tests/test_ref_pin.py); the Eigen calls inside are restated -- see trackdlo_oracle.cpp header.
"""
from .oracle import (build, lib, cpd_lle, tracking_step, traverse_euclidean, lle_H, visibility, tracking_error,  # noqa: F401
                     CpdParams, TrackParams)

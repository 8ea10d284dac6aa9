"""CPU oracle for the TrackDLO registration path -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; the product package (trackdlo_b200) never imports it.  PARITY UNPINNED (see
trackdlo_oracle.cpp header): the reference has no golden vectors and is not buildable here.
"""
from .oracle import (build, lib, cpd_lle, tracking_step, traverse_euclidean, lle_H, visibility, tracking_error,  # noqa: F401
                     CpdParams, TrackParams)

"""CPU oracle for the TubeR forward path -- test infrastructure, never imported by the product."""

"""CPU oracle for the GenS hot path -- TEST INFRASTRUCTURE, never imported by gens_b200/."""

"""Action ids of the MAPF environment (reference: src/environments/actions.py:1-5)."""
NO_OP, UP, RIGHT, DOWN, LEFT = range(5)

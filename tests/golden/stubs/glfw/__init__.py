"""``wrappers.py:2`` imports glfw for PlayWrapper (GUI, out of scope)."""

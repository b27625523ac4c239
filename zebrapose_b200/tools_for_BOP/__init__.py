"""Drop-in for the reference package `tools_for_BOP` (only the result writer is on the pose path's output side)."""

"""``matplotlib`` is imported by the reference agent for its end-of-training plot only."""

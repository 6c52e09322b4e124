class Box(object):
    """Never instantiated by the point envs; only needs to exist for isinstance()."""

"""No-op stand-in for matplotlib (absent in the image); only used so the reference imports."""

def sym(x):
    return 0.5 * (x.transpose(-1, -2) + x)

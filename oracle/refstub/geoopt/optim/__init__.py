class RiemannianSGD:  # placeholder; the reference only names it in train.py
    pass


class RiemannianAdam:
    pass

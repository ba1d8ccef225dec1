class BCOO:  # placeholders so `import jax.experimental.sparse as jsp` succeeds
    pass


class BCSR:
    pass

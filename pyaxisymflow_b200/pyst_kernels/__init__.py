"""Mirror of the reference package ``pyaxisymflow.pyst_kernels`` (same module and callable names)."""

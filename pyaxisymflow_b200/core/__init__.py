"""Mirror of the reference package ``pyaxisymflow.core`` (same module and callable names)."""

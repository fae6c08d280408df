"""B200-native hot path of 3D-DyCorePlanet: Boussinesq FE assembly + CSR SpMV behind a C ABI.

The directory name is not a Python identifier; import it through the root-level shim `dycore_b200`.
"""

"""Import shims that give this package the module paths the reference's main.py / engine.py
import (`models`, `criterion`, `third_party_pointnet2.pointnet2.*`, `pointnet2._ext`).
Put this directory first on sys.path (see launch.py / INTEGRATION.md)."""

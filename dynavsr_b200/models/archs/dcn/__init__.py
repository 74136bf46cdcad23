"""Mirror of codes/models/archs/dcn/__init__.py:1-7 (DCNv2 exports; DCNv1 is not on the hot path)."""
from .deform_conv import (ModulatedDeformConv, ModulatedDeformConvFunction, ModulatedDeformConvPack,
                          modulated_deform_conv)

__all__ = ['ModulatedDeformConv', 'ModulatedDeformConvPack', 'ModulatedDeformConvFunction',
           'modulated_deform_conv']

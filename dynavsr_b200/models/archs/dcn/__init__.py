"""Mirror of codes/models/archs/dcn/__init__.py:1-7 (same export list)."""
from .deform_conv import (DeformConv, DeformConvPack, ModulatedDeformConv, ModulatedDeformConvPack, deform_conv,
                          modulated_deform_conv)
from .deform_conv import DeformConvFunction, ModulatedDeformConvFunction  # noqa: F401

__all__ = ['DeformConv', 'DeformConvPack', 'ModulatedDeformConv', 'ModulatedDeformConvPack', 'deform_conv',
           'modulated_deform_conv']

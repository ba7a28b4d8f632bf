"""Physical constants of the coupling path, plain SI floats.

Same values as the reference's unit-carrying constants (splib/sputils.py:14-19).
"""
pref0 = 1.0e5    # Pa, reference pressure            (sputils.py:14)
rd = 287.04      # J/kg/K, gas constant of dry air   (sputils.py:15)
rv = 461.5       # J/kg/K, gas constant of vapour    (sputils.py:16)
cp = 1004.0      # J/kg/K, heat capacity of dry air  (sputils.py:17)
rlv = 2.53e6     # J/kg, latent heat of vaporisation (sputils.py:18)
grav = 9.81      # m/s^2                             (sputils.py:19)

# Field order of the five slab-averaged LES volumes on the path (spcpl.py:748-755)
LES_FIELDS = ("THL", "QT", "QL", "U", "V")
F_THL, F_QT, F_QL, F_U, F_V = range(5)

# reference names (spcpl.py:32-33)
gcm_vars = ["U", "V", "T", "SH", "QL", "QI", "Pfull", "Phalf", "A", "Zgfull", "Zghalf"]
surf_vars = ["Z0M", "Z0H", "QLflux", "QIflux", "SHflux", "TLflux", "TSflux"]

# order of the packed GCM tendency block [ncol][7][nlev] (spcpl.py:518-526)
TENDENCIES = ("f_T", "f_SH", "f_QL", "f_QI", "f_U", "f_V", "f_A")

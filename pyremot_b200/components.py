"""Component property database of the N1/N2 path, in table form for the device code.

Same species, order and values as the reference's database (`PyREMOT/data/componentData.py:11-22`
MW, `:72-86` dHf25, Cp expression strings `:119, 145, 171, 197, 223, 249, ...`; viscosity parameters
`PyREMOT/data/dataGasViscosity.py:8-141`).  The reference keeps Cp as python expression strings that it
``eval``s on every RHS call (`PyREMOT/docs/rmtThermo.py:37`, 65 % of its run time); here they are the
polynomial coefficients ``a + b T + c T^2 + d T^3`` that are baked into the NVRTC-compiled kernels.
DME's viscosity is an expression string in the reference (`dataGasViscosity.py:133`,
``2.68e-7*T**0.3975/(1+534/T)``); it is the eq1 form with A=0.268, D=0.
"""
from __future__ import annotations

# symbol -> (MW [g/mol], Cp [a, b, c, d] J/mol/K, dHf25 [kJ/mol], viscosity eq1 [A, B, C, D])
COMPONENTS = {
    "CO2":   (44.01, (22.243, 5.98e-02, -3.50e-05, 7.46e-09), -393.51, (4.719875, 0.373279, 512.686300, -6119.961)),
    "H2":    (2.0,   (26.879, 4.35e-03, -3.30e-07, 0.0), 0.0, (0.169104, 0.692485, -7.634394, 467.120)),
    "CH3OH": (32.04, (19.038, 9.15e-02, -1.22e-05, -8.03e-09), -200.7, (0.477915, 0.641076, 284.838034, -3230.713)),
    "H2O":   (18.01, (29.163, 1.45e-02, -2.02e-06, 0.0), -241.820, (0.501246, 0.709247, 869.465599, -90063.891)),
    "CO":    (28.01, (27.113, 6.55e-03, -1.00e-06, 0.0), -110.53, (0.734306, 0.588574, 52.318660, 1018.822)),
    "DME":   (46.07, (19.8, 0.17, -5.66e-5, 0.0), -184.1, (0.268, 0.3975, 534.0, 0.0)),
    "N2":    (28.0,  (28.883, -1.57e-03, 8.08e-06, -2.87e-09), 0.0, (0.847662, 0.574033, 75.437536, 56.771)),
    "CH4":   (16.04, (19.875, 5.021e-02, 1.268e-05, -11.004e-09), -74.90, (1.119178, 0.493234, 214.627200, -3952.087)),
    "C2H4":  (28.05, (3.950, 15.628e-02, -8.339e-05, 17.657e-09), 52.32, (1.503552, 0.456140, 288.342422, 73.362)),
    "C3H6":  (42.08, (3.151, 23.812e-02, -12.176e-05, 24.603e-09), 20.4, (0.876767, 0.520871, 293.618650, -182.857)),
    "C3H8":  (44.1,  (-4.042, 30.456e-02, -15.711e-05, 31.716e-09), -103.9, (0.173966, 0.734798, 143.207060, -7147.859)),
    "C4H10": (58.12, (-7.908, 41.573e-02, -22.992e-05, 49.875e-09), -126.2, (0.075828, 0.837082, 67618677.0, -2141.762)),
}

componentSymbolList = tuple(COMPONENTS)  # `PyREMOT/data/componentData.py:435`


def MW(s):
    return COMPONENTS[s][0]


def cp_coef(s):
    return COMPONENTS[s][1]


def dHf25(s):
    return COMPONENTS[s][2]


def vis_coef(s):
    return COMPONENTS[s][3]

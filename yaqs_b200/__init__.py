"""yaqs_b200 -- B200-native batched Tensor Jump Method behind the MQT YAQS public surface.

Drop-in names (reference: src/mqt/yaqs/__init__.py:39-59): Simulator, State, Hamiltonian, NoiseModel, AnalogSimParams,
DigitalSimParams, Observable, Result, SimulationProgram, MPS, MPO, plus the pre-0.6 aliases StrongSimParams / WeakSimParams.
Importing the package never touches the GPU; the CUDA library is loaded on first use and there is no CPU fallback.
"""
from .circuits import (Circuit, compile_circuit, create_1d_fermi_hubbard_circuit, create_2d_fermi_hubbard_circuit, create_2d_heisenberg_circuit, create_2d_ising_circuit,
                       create_heisenberg_circuit, create_ising_circuit, nearest_neighbour_random_circuit)
from .containers import MPO, MPS, Hamiltonian, Result, State
from .gates import GateLibrary, X, XX, Y, YY, Z, ZZ
from .noise import NoiseModel
from .params import AnalogSimParams, DigitalSimParams, Observable, StrongSimParams, WeakSimParams
from .program import SimulationProgram
from .simulator import Simulator

__all__ = ["MPO", "MPS", "AnalogSimParams", "Circuit", "compile_circuit", "create_ising_circuit", "create_heisenberg_circuit",
           "create_2d_ising_circuit", "create_2d_heisenberg_circuit", "create_1d_fermi_hubbard_circuit", "create_2d_fermi_hubbard_circuit", "nearest_neighbour_random_circuit", "DigitalSimParams", "GateLibrary", "Hamiltonian", "NoiseModel", "Observable", "Result",
           "SimulationProgram", "Simulator", "State", "StrongSimParams", "WeakSimParams", "X", "XX", "Y", "YY", "Z", "ZZ"]
__version__ = "0.1.0"

"""Digital TJM entry (digital/digital_tjm.py:636-749).  Built after the analog path."""


def run_digital(simulator, initial_state, operator, sim_params, noise_model, result):
    raise NotImplementedError("digital TJM on the B200 path is not built yet (SURVEY.md section 7, step 12); "
                              "the analog TJM (AnalogSimParams) path is complete")

"""TEST INFRASTRUCTURE — re-export of the synthetic weight / condition generators (tools/synthetic.py) under the
name the parity tests use."""
from tools.synthetic import *  # noqa: F401,F403
from tools.synthetic import (DAC_CONFIG, DAC_TINY, MODEL_CONFIGS, clip_lengths, dac_param_specs, dit_param_specs,  # noqa: F401
                             model_config, synth_conditions, synth_dac_state_dict, synth_dit_state_dict)

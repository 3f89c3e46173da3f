from peclr_b200.model_utils import (get_encoder_state_dict, get_latest_checkpoint, get_wrapper_model,  # noqa: F401
                                    vanila_contrastive_loss)

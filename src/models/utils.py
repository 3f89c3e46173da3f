from peclr_b200.model_utils import (get_encoder_state_dict, get_latest_checkpoint, get_rotation_2D_matrix,  # noqa: F401
                                    get_wrapper_model, rotate_encoding, translate_encodings, translate_encodings2,
                                    vanila_contrastive_loss)

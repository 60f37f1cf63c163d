/* TEST INFRASTRUCTURE — CPU oracle for the MoPA-RL hot path.  Not part of the product.
 *
 * Contact generation for the physics-step oracle (placeholder: contact rows follow). */
int orc_contact_rows_placeholder(void) { return 0; }

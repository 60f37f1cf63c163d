/* TEST INFRASTRUCTURE — placeholder, dynamics oracle follows. */

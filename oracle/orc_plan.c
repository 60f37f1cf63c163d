/* TEST INFRASTRUCTURE — placeholder, RRT-Connect oracle follows. */

/* placeholder, BA oracle follows */

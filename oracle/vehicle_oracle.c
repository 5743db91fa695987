/* vehicle_oracle.c -- placeholder, filled in with the tick restatement */
int oracle_vehicle_placeholder(void) { return 0; }
